"""Scratch: where does the persistent halo conv kernel lose time?  Times a shape with weight streaming and/or halo loads disabled
(tc_variant debug bits; results are wrong by construction)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sma_b200 as S

def run(B, Cin, H, Cout, k, pad):
    x = torch.randn(B, H, H, Cin, device='cuda'); w = torch.randn(Cout, Cin, k, k, device='cuda') * (Cin * k * k) ** -0.5
    cw = S.ops.pack_conv(w, torch.randn(Cout, device='cuda'))
    res = []
    for ts, f16 in ((True, True), (False, True)):
        for fast in (False, True):
            for dbg in (0, 2, 4, 8, 14):
                S.ops.USE_TS, S.ops.USE_F16, S.ops.TC_VARIANT = ts, f16, dbg
                y = S.ops.conv2d(x, cw, pad=pad, fast=fast)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                for _ in range(5): S.ops.conv2d(x, cw, pad=pad, out=y, fast=fast)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                prof = ''
                if ts:
                    import ctypes
                    buf = (ctypes.c_longlong * 8)()
                    S._lib.load().sma_debug_conv_ts_prof(ctypes.cast(buf, ctypes.c_void_p))
                    mmas = (B * ((H + 15) // 16) * ((H + 7) // 8) * (1 if Cout <= 64 else (Cout + 127) // 128) / 148.0) * (Cin // 64) * k * k * 4 * ((2 if Cout <= 64 else 3) if not fast else 1) * (0.5 if (k > 1 and Cin * k * k // 64 * (1 if Cout <= 64 else 2) * 32 > 288) else 1.0)
                    prof = f' | CTA0 {buf[0]/1e3:.0f} kcyc {buf[1]/1e3:.0f} us -> {buf[0]/max(buf[1],1)*1e3:.0f} MHz, {buf[0]/mmas:.0f} cyc/MMA'
                res.append(f"{'ts' if ts else 'ss'}{'x1' if fast else 'x3'} dbg{dbg}: {ms:.3f} ms {2.0*B*H*H*Cin*k*k*Cout/ms/1e9:.0f} TF" + prof)
    S.ops.USE_TS, S.ops.USE_F16, S.ops.TC_VARIANT = True, True, 0
    print(f'B{B} Cin{Cin} H{H} Cout{Cout} k{k}\n   ' + '\n   '.join(res), flush=True)

for a in [(64, 64, 256, 64, 3, 1), (64, 128, 128, 128, 3, 1), (64, 256, 64, 256, 3, 1), (64, 128, 256, 64, 3, 1), (64, 256, 32, 512, 3, 1)]:
    run(*a)
