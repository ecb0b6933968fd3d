"""Scratch: one conv shape, a few launches (for ncu).  usage: one_conv.py B Cin H Cout k [pre] [res] [fast] [up] [relu]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sma_b200 as S
B, Cin, H, Cout, k = map(int, sys.argv[1:6])
pre, res, fast, up = 'pre' in sys.argv, 'res' in sys.argv, 'fast' in sys.argv, 'up' in sys.argv
act = 'relu' if 'relu' in sys.argv else 'none'
x = torch.randn(B, H, H, Cin, device='cuda'); w = torch.randn(Cout, Cin, k, k, device='cuda') * (Cin * k * k) ** -0.5
cw = S.ops.pack_conv(w, torch.randn(Cout, device='cuda'))
prek = (torch.ones(B, Cin, device='cuda'), torch.zeros(B, Cin, device='cuda'), 'swish') if pre else None
r = torch.randn(B, H * (2 if up else 1), H * (2 if up else 1), Cout, device='cuda') if res else None
for _ in range(3):
    y = S.ops.conv2d(x, cw, pad=k // 2, pre=prek, res=r, fast=fast, upsample2=up, act=act)
torch.cuda.synchronize()
