"""Scratch: what in the RES = 1 epilogue costs time?  tc_variant bit 13: no stores (loads still waited for), bit 14: residual values unused, bit 3: neither."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sma_b200 as S
def t(B, Cin, H, Cout, k, pad, res, var):
    x = torch.randn(B, H, H, Cin, device='cuda'); w = torch.randn(Cout, Cin, k, k, device='cuda') * (Cin * k * k) ** -0.5
    cw = S.ops.pack_conv(w, torch.randn(Cout, device='cuda'))
    r = torch.randn(B, H, H, Cout, device='cuda') if res else None
    S.ops.TC_VARIANT = var
    y = S.ops.conv2d(x, cw, pad=pad, res=r)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): S.ops.conv2d(x, cw, pad=pad, out=y, res=r)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
for shp in [(64, 64, 256, 64, 3, 1, True), (64, 256, 32, 256, 1, 0, True), (64, 128, 128, 128, 3, 1, True)]:
    for var, name in [(0, 'all on'), (8192, 'no stores'), (16384, 'residual unused'), (8192 + 16384, 'no stores, residual unused'), (8, 'no epilogue memory ops')]:
        print(shp, f'{name:28s}', '%.3f ms' % t(*shp, var), flush=True)
