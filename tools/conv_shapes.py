"""Scratch: time the main convolution shapes of one 64-frame step (all features on), one line per shape; used to A/B kernel variants
(SMA_B200_LIB=<variant .so>)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sma_b200 as S
SHAPES = [  # B, Cin, H, Cout, k, pre, res, fast, weight in the step (launches)
    (64, 64, 256, 64, 3, True, True, False, 4), (64, 64, 256, 64, 3, False, True, False, 3), (64, 128, 256, 64, 3, True, False, False, 2),
    (64, 128, 128, 128, 3, True, True, False, 5), (64, 128, 128, 128, 3, False, False, False, 4), (64, 256, 32, 512, 3, False, False, False, 8),
    (64, 512, 32, 256, 3, False, True, False, 8), (64, 256, 32, 256, 1, False, True, False, 32), (64, 128, 256, 64, 1, False, False, False, 2),
    (64, 64, 256, 192, 1, False, False, True, 1), (64, 256, 32, 256, 3, True, True, False, 8), (64, 256, 64, 256, 3, False, False, True, 4),
    (64, 160, 64, 126, 3, False, False, True, 4), (64, 64, 256, 128, 3, False, False, False, 1), (64, 128, 64, 128, 3, True, True, False, 7),
    (64, 64, 256, 3, 3, True, False, False, 1), (64, 256, 32, 4096, 1, False, False, False, 1)]
S.ops.TC_VARIANT = int(os.environ.get('TCV', '0'))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
tot = 0.0
rows = []
for (B, Cin, H, Cout, k, pre, res, fast, wgt) in SHAPES:
    x = torch.randn(B, H, H, Cin, device='cuda'); w = torch.randn(Cout, Cin, k, k, device='cuda') * (Cin * k * k) ** -0.5
    cw = S.ops.pack_conv(w, torch.randn(Cout, device='cuda'))
    prek = (torch.ones(B, Cin, device='cuda'), torch.zeros(B, Cin, device='cuda'), 'swish') if pre else None
    r = torch.randn(B, H, H, Cout, device='cuda') if res else None
    y = S.ops.conv2d(x, cw, pad=k // 2, fast=fast, pre=prek, res=r)
    ts = []
    for _ in range(4):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S.ops.conv2d(x, cw, pad=k // 2, out=y, fast=fast, pre=prek, res=r); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[1]
    tot += ms * wgt
    rows.append(f'{ms:.3f}')
print('TCV', S.ops.TC_VARIANT, os.environ.get('SMA_B200_LIB', 'default').split('/')[-1], f'weighted {tot:.2f} ms |', ' '.join(rows), flush=True)
