"""Launch the small HBM-bound kernels at the 64-frame step's largest shapes, for
`ncu --metrics gpu__time_duration.sum -k regex:"tapsum|layernorm"` (cold-cache per-launch durations, comparable to the launch list's)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sma_b200 as S
ops = S.ops
g = torch.Generator(device='cuda').manual_seed(0)
for B, H in [(64, 256), (64, 64)]:
    P = torch.randn(B, H, H, 32, device='cuda', generator=g)
    for _ in range(2):
        o = ops.conv_tapsum(P, torch.zeros(3, device='cuda'), 3, 3, 1)
x = torch.randn(64, 1024, 32, device='cuda', generator=g)
ga, be, pos = torch.ones(32, device='cuda'), torch.zeros(32, device='cuda'), torch.randn(1024, 32, device='cuda', generator=g)
for _ in range(2):
    y, yq = ops.layernorm(x, ga, be, pos)
    y2, _ = ops.layernorm(x, ga, be, None)
torch.cuda.synchronize()
print('ok')
