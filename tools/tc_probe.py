"""Scratch probe: accuracy and speed of the tcgen05 conv kernel vs the exact CUDA-core kernel (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import sma_b200 as S

def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale

def run(B, Cin, H, Cout, k, pad, reps=5):
    x = rnd(B, Cin, H, H, seed=1); w = rnd(Cout, Cin, k, k, seed=2, scale=(Cin*k*k) ** -0.5); b = rnd(Cout, seed=3, scale=0.1)
    ref = F.conv2d(x.double().cuda(), w.double().cuda(), b.double().cuda(), padding=pad).permute(0, 2, 3, 1)
    cw = S.ops.pack_conv(w.cuda(), b.cuda()); xh = x.permute(0, 2, 3, 1).contiguous().cuda()
    out = {}
    for name, kw in (('exact', dict(exact=True)), ('v1x3', {}), ('v2x3', {}), ('v2x1', dict(fast=True))):
        S.ops.TC_VARIANT = 1 if name.startswith('v1') else 0
        y = S.ops.conv2d(xh, cw, pad=pad, **kw); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): S.ops.conv2d(xh, cw, pad=pad, out=y, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 2.0 * B * H * H * Cin * k * k * Cout
        out[name] = (float((y.double() - ref).abs().max()), float((y.double() - ref).abs().mean()), ms, fl / ms / 1e9)
    print(f'B{B} Cin{Cin} H{H} Cout{Cout} k{k}: refmax {float(ref.abs().max()):.2f} | ' + ' | '.join(f'{n}: max {v[0]:.2e} mean {v[1]:.2e} {v[2]:.3f} ms {v[3]:.1f} TF' for n, v in out.items()), flush=True)

for a in [(1, 128, 64, 17, 7, 3), (1, 128, 64, 32, 7, 3), (1, 128, 64, 128, 7, 3), (1, 128, 64, 16, 3, 1), (1, 128, 64, 32, 3, 1), (1, 128, 64, 48, 3, 1), (1, 128, 64, 64, 3, 1),
          (16, 64, 256, 64, 3, 1), (16, 128, 128, 128, 3, 1), (16, 256, 64, 256, 3, 1), (16, 256, 32, 512, 3, 1), (16, 512, 32, 256, 3, 1),
          (16, 256, 32, 768, 1, 0), (16, 128, 256, 64, 3, 1), (16, 1024, 4, 1024, 3, 1), (16, 64, 256, 128, 3, 1)]:
    run(*a)
