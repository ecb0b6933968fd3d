"""Scratch: where does the persistent halo conv kernel (conv_tc2, fp16 split) lose time?  Times a shape with the weight ring traffic,
the halo producers and / or the epilogue stores disabled (tc_variant bits 1-3; results are wrong by construction)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sma_b200 as S

def run(B, Cin, H, Cout, k, pad, pre=False, res=False):
    x = torch.randn(B, H, H, Cin, device='cuda'); w = torch.randn(Cout, Cin, k, k, device='cuda') * (Cin * k * k) ** -0.5
    cw = S.ops.pack_conv(w, torch.randn(Cout, device='cuda'))
    prek = (torch.ones(B, Cin, device='cuda'), torch.zeros(B, Cin, device='cuda'), 'swish') if pre else None
    r = torch.randn(B, H, H, Cout, device='cuda') if res else None
    out = []
    for fast in (False,):
        for dbg, name in ((0, 'all on'), (2, 'no weights'), (4, 'no halo'), (8, 'no epilogue'), (6, 'no weights, no halo'), (14, 'MMA only')):
            S.ops.TC_VARIANT = dbg
            y = S.ops.conv2d(x, cw, pad=pad, fast=fast, pre=prek, res=r)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(5): S.ops.conv2d(x, cw, pad=pad, out=y, fast=fast, pre=prek, res=r)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            out.append(f"{'x1' if fast else 'x3'} {name:20s}: {ms:.3f} ms {2.0*B*H*H*Cin*k*k*Cout/ms/1e9:.0f} TF")
    S.ops.TC_VARIANT = 0
    print(f'B{B} Cin{Cin} H{H} Cout{Cout} k{k} pre={pre} res={res}\n   ' + '\n   '.join(out), flush=True)

for a in [(64, 64, 256, 64, 3, 1, True, True), (64, 64, 256, 64, 3, 1, False, True), (64, 128, 128, 128, 3, 1, True, True), (64, 128, 256, 64, 3, 1, True), (64, 128, 256, 64, 1, 0), (64, 64, 256, 192, 1, 0), (64, 256, 32, 256, 1, 0, False, True)]:
    run(*a)
