"""Scratch: is the shared-memory-operand halo kernel bound by its epilogue?  (tc_variant bit 3 = no epilogue stores; results invalid)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sma_b200 as S
S.ops.USE_TS = False
def run(B, Cin, H, Cout, k, pad, res=False, pre=False):
    x = torch.randn(B, H, H, Cin, device='cuda'); w = torch.randn(Cout, Cin, k, k, device='cuda') * (Cin * k * k) ** -0.5
    cw = S.ops.pack_conv(w, torch.randn(Cout, device='cuda'))
    kw = dict(pad=pad)
    if res: kw['res'] = torch.randn(B, H, H, Cout, device='cuda')
    if pre: kw['pre'] = (torch.rand(B, Cin, device='cuda') + 0.5, torch.randn(B, Cin, device='cuda') * 0.1, 'swish')
    out = []
    for dbg in (0, 8, 128, 136):
        S.ops.TC_VARIANT = dbg
        y = S.ops.conv2d(x, cw, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5): S.ops.conv2d(x, cw, out=y, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out.append(f'variant {dbg}: {ms:.3f} ms {2.0*B*H*H*Cin*k*k*Cout/ms/1e9:.0f} TF')
    S.ops.TC_VARIANT = 0
    print(f'B{B} Cin{Cin} H{H} Cout{Cout} k{k} res{int(res)} pre{int(pre)}: ' + ' | '.join(out), flush=True)
run(64, 64, 256, 64, 3, 1); run(64, 64, 256, 64, 3, 1, res=True, pre=True); run(64, 128, 256, 64, 3, 1); run(64, 128, 128, 128, 3, 1); run(64, 256, 32, 256, 1, 0); run(64, 128, 256, 64, 1, 0)
