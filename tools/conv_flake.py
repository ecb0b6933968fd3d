"""Run-to-run equality of single convolutions per kernel family: python tools/conv_flake.py [iterations]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sma_b200 as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = torch.Generator().manual_seed(1)
CASES = [(2, 256, 32, 32, 768, 1, 0, {}), (1, 256, 32, 32, 256, 3, 1, {'res': True}), (2, 128, 16, 16, 256, 3, 1, {}), (64, 256, 32, 32, 768, 1, 0, {}), (8, 64, 128, 128, 64, 3, 1, {'res': True})]
for B, Cin, H, W, Cout, k, pad, ex in CASES:
    x = torch.randn(B, H, W, Cin, generator=g).cuda()
    w = (torch.randn(Cout, Cin, k, k, generator=g) * (Cin * k * k) ** -0.5).cuda()
    b = (torch.randn(Cout, generator=g) * 0.1).cuda()
    res = torch.randn(B, H, W, Cout, generator=g).cuda() if ex.get('res') else None
    cw = S.ops.pack_conv(w, b)
    for mode in ('default', 'f16x3', 'ts', 'ts-stream', 'tf32x3', 'gather'):
        saved = (S.ops.TC_VARIANT, S.ops.USE_F16, S.ops.USE_TS)
        S.ops.TC_VARIANT = {'gather': 1, 'ts': 32, 'ts-stream': 48}.get(mode, 0)
        S.ops.USE_F16 = mode in ('f16x3', 'ts', 'ts-stream', 'default')
        S.ops.USE_TS = mode in ('ts', 'ts-stream', 'default')
        try:
            first = S.ops.conv2d(x, cw, pad=pad, res=res).clone()
            kern = S.ops.LAST_CONV_KERNEL
            bad = 0
            for _ in range(n):
                y = S.ops.conv2d(x, cw, pad=pad, res=res)
                if not torch.equal(y, first):
                    bad += 1
                    if bad == 1:
                        d = (y != first)
                        print('   first mismatch: %d values, max diff %.3e, at %s' % (int(d.sum()), float((y - first).abs().max()), d.nonzero()[0].tolist()))
        finally:
            S.ops.TC_VARIANT, S.ops.USE_F16, S.ops.USE_TS = saved
        print(f'B{B} Cin{Cin} {H}x{W} Cout{Cout} k{k} {mode:10s} kernel {kern}: {bad} of {n} launches differ', flush=True)
