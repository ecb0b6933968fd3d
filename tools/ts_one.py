"""Scratch: launch one conv shape on the tensor-memory-operand kernel a few times (for ncu).  args: B Cin H Cout k pad dbg"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sma_b200 as S
B, Cin, H, Cout, k, pad, dbg = [int(a) for a in sys.argv[1:8]]
x = torch.randn(B, H, H, Cin, device='cuda'); w = torch.randn(Cout, Cin, k, k, device='cuda') * (Cin * k * k) ** -0.5
cw = S.ops.pack_conv(w, torch.randn(Cout, device='cuda'))
S.ops.TC_VARIANT = dbg
y = S.ops.conv2d(x, cw, pad=pad)
torch.cuda.synchronize()
torch.cuda.profiler.start()
S.ops.conv2d(x, cw, pad=pad, out=y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('kernel', S.ops.LAST_CONV_KERNEL)
