import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sma_b200 as S
B, Cin, H, Cout, k, pad = [int(a) for a in sys.argv[1:7]]
fast = len(sys.argv) > 7 and sys.argv[7] == 'fast'
x = torch.randn(B, H, H, Cin, device='cuda'); w = torch.randn(Cout, Cin, k, k, device='cuda') * (Cin * k * k) ** -0.5; b = torch.randn(Cout, device='cuda')
cw = S.ops.pack_conv(w, b)
y = S.ops.conv2d(x, cw, pad=pad, fast=fast)
for _ in range(3): S.ops.conv2d(x, cw, pad=pad, out=y, fast=fast)
torch.cuda.synchronize()
