"""Launch the GroupNorm finalize (through a convolution with fused statistics) and the tap gather-sum at the 64-frame step's largest shapes, for
`ncu --metrics gpu__time_duration.sum -k regex:"tapsum|gn_finalize"` (cold-cache per-launch durations, comparable to the launch list's)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sma_b200 as S
ops = S.ops
g = torch.Generator(device='cuda').manual_seed(0)
for B, H, C in [(64, 256, 64), (64, 128, 128), (1, 256, 64), (64, 64, 256)]:
    x = torch.randn(B, H, H, C, device='cuda', generator=g)
    cw = ops.pack_conv(torch.randn(C, C, 3, 3, device='cuda', generator=g) * (9 * C) ** -0.5, torch.zeros(C, device='cuda'))
    gamma, beta = torch.ones(C, device='cuda'), torch.zeros(C, device='cuda')
    for _ in range(2):
        y, (sc, sh) = ops.conv2d(x, cw, pad=1, gn=(gamma, beta))
    sc2, sh2 = ops.groupnorm_stats(y, gamma, beta, 32, 1e-6)
    print(B, H, C, 'finalize vs standalone pass: scale', float((sc - sc2).abs().max() / sc2.abs().max()), 'shift', float((sh - sh2).abs().max()), flush=True)
    del x, y
for B, H in [(64, 256), (64, 64)]:
    P = torch.randn(B, H, H, 32, device='cuda', generator=g)
    for _ in range(2):
        o = ops.conv_tapsum(P, torch.zeros(3, device='cuda'), 3, 3, 1)
torch.cuda.synchronize()
print('ok')
