import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sma_b200 as S
B, L, Skv, heads, D = [int(a) for a in sys.argv[1:6]]
E = heads * D
q = torch.randn(B, L, E, device='cuda'); k = torch.randn(B, Skv, E, device='cuda'); v = torch.randn(B, Skv, E, device='cuda')
o = S.ops.mha(q, k, v, heads)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5): S.ops.mha(q, k, v, heads, out=o)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(sys.argv[1:], f'{ms:.3f} ms  {4.0*B*L*Skv*E/ms/1e9:.1f} TF')
