"""config 4 fed NH blocks per call (multi-hop reuse), a few calls: the command ncu wraps for the k_cmac_mh2 capture"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from hisstools_library_b200.convolve import _Engine
nh = int(sys.argv[1]) if len(sys.argv) > 1 else 4
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ins = outs = 64; B = 4096; P = 64; taps = B * P
dev = torch.device("cuda", 0)
eng = _Engine(np.float32, 1, ins, outs, 2 * B, taps, 0, 0, 0)
eng.set_reset_offset(0)
ir = torch.randn(taps, device=dev) * torch.exp(-6.9 * torch.arange(taps, device=dev) / taps)
for o in range(outs):
    for i in range(ins):
        eng.set_ir_device(0, i, o, ir.data_ptr(), taps)
x = torch.rand(ins, nh * B, device=dev) * 2 - 1
y = torch.zeros(outs, nh * B, device=dev)
st = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(st):
    for k in range(2):
        eng.process_device(x.data_ptr(), nh * B, y.data_ptr(), nh * B, nh * B, False, st.cuda_stream)
    e0.record(st)
    for k in range(calls):
        eng.process_device(x.data_ptr(), nh * B, y.data_ptr(), nh * B, nh * B, False, st.cuda_stream)
    e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / calls
print("nh %d: %.3f ms per call, %.1f M output-samples/s" % (nh, ms, outs * nh * B / ms / 1e3))
