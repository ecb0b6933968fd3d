"""Scratch: is the fp16-split tensor-core conv error a systematic truncation bias of the TMEM accumulation?  signed relative error vs K."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import sma_b200 as S
torch.manual_seed(0)
for (Cin, Cout, k, pos) in [(64, 64, 3, False), (128, 128, 3, False), (256, 256, 3, False), (512, 256, 3, False), (256, 256, 1, False), (128, 128, 3, True), (512, 256, 3, True)]:
    x = torch.randn(2, Cin, 32, 32)
    w = torch.randn(Cout, Cin, k, k) * (Cin * k * k) ** -0.5
    if pos:                      # all-positive terms: partial sums grow monotonically (worst case for truncation)
        x, w = x.abs(), w.abs()
    ref = F.conv2d(x.double(), w.double(), None, padding=k // 2)
    xc = x.permute(0, 2, 3, 1).contiguous().cuda()
    cw = S.ops.pack_conv(w.cuda(), None)
    for mode in ('f16x3', 'exact'):
        y = S.ops.conv2d(xc, cw, pad=k // 2, exact=(mode == 'exact')).permute(0, 3, 1, 2).cpu().double()
        e = y - ref
        big = ref.abs() > ref.abs().mean()
        rel = (e[big] / ref[big])
        n_adds = k * k * Cin // 16
        print(f'Cin {Cin} Cout {Cout} k {k} pos {pos} {mode}: kernel {S.ops.LAST_CONV_KERNEL} adds/out {n_adds} mean signed rel {float(rel.mean()):+.3e} rms rel {float(rel.pow(2).mean().sqrt()):.3e} '
              f'| signed*sign(ref) abs-mean {float((e * ref.sign()).mean()):+.3e} rms abs {float(e.pow(2).mean().sqrt()):.3e}')
