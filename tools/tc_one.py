import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sma_b200 as S
B, Cin, H, Cout, k, pad = [int(a) for a in sys.argv[1:7]]
flags = sys.argv[7:]
x = torch.randn(B, H, H, Cin, device='cuda'); w = torch.randn(Cout, Cin, k, k, device='cuda') * (Cin * k * k) ** -0.5; b = torch.randn(Cout, device='cuda')
cw = S.ops.pack_conv(w, b)
kw = dict(pad=pad, fast='fast' in flags)
if 'pre' in flags: kw['pre'] = (torch.rand(B, Cin, device='cuda') + 0.5, torch.randn(B, Cin, device='cuda') * 0.1, 'swish')
if 'res' in flags: kw['res'] = torch.randn(B, H, H, Cout, device='cuda')
y = S.ops.conv2d(x, cw, **kw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5): S.ops.conv2d(x, cw, out=y, **kw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(sys.argv[1:], f'{ms:.3f} ms  {2.0*B*H*H*Cin*k*k*Cout/ms/1e9:.1f} TF')
