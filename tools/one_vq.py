"""Scratch: the VQ lookup on one batch of appearance tokens (for ncu).  usage: one_vq.py [N] [E] [n_codes]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sma_b200 as S
N, E, n = (int(a) for a in (sys.argv[1:4] + ['65536', '256', '1024'][len(sys.argv) - 1:]))
z = torch.randn(N, E, device='cuda'); cb = torch.randn(1024, E, device='cuda')
for _ in range(3):
    idx, zq, md = S.ops.vq_lookup(z, cb, n)
    st, loss, _ = S.ops.vq_quantize(z, cb, n, 0.25)
torch.cuda.synchronize()
print(int(idx[0]), float(loss))
