"""Scratch: one conv shape, a few launches (for ncu).  usage: one_conv.py B Cin H Cout k [pre] [res]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sma_b200 as S
B, Cin, H, Cout, k = map(int, sys.argv[1:6])
pre, res = 'pre' in sys.argv, 'res' in sys.argv
x = torch.randn(B, H, H, Cin, device='cuda'); w = torch.randn(Cout, Cin, k, k, device='cuda') * (Cin * k * k) ** -0.5
cw = S.ops.pack_conv(w, torch.randn(Cout, device='cuda'))
prek = (torch.ones(B, Cin, device='cuda'), torch.zeros(B, Cin, device='cuda'), 'swish') if pre else None
r = torch.randn(B, H, H, Cout, device='cuda') if res else None
for _ in range(3):
    y = S.ops.conv2d(x, cw, pad=k // 2, pre=prek, res=r)
torch.cuda.synchronize()
