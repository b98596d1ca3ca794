"""hop period of small matrices, automatic (fused) against serial schedule: python tools/small_matrix_probe.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hisstools_library_b200.convolve import _Engine

dev = torch.device("cuda", 0)
st = torch.cuda.Stream()
for ins, outs, taps, B in ((1, 1, 4096, 512), (2, 2, 4096, 512), (2, 2, 65536, 1024), (4, 4, 16384, 256), (2, 8, 8192, 128)):
    row = []
    for sched in (None, False):
        e = _Engine(np.float32, 1, ins, outs, 2 * B, taps, 0, 0, 0)
        e.set_schedule(sched)
        e.set_reset_offset(0)
        ir = torch.randn(taps, device=dev)
        for o in range(outs):
            for i in range(ins):
                e.set_ir_device(0, i, o, ir.data_ptr(), taps)
        x = torch.rand(ins, B, device=dev)
        y = torch.zeros(outs, B, device=dev)
        for _ in range(10):
            e.process_device(x.data_ptr(), B, y.data_ptr(), B, B, False, st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(200):
            e.process_device(x.data_ptr(), B, y.data_ptr(), B, B, False, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        row.append("%s %.1f us/hop" % (e.schedule, e0.elapsed_time(e1) * 1e3 / 200))
        e.close()
    print("%d in x %d out, %6d taps, hop %4d: %s" % (ins, outs, taps, B, "  |  ".join(row)))
