"""Per-call wall-clock of hb_conv_process (host pointers) at a bench workload: pipelined vs synchronous."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hisstools_library_b200.convolve import _Engine

def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
    cfg = {"c4": (64, 64, 1, 262144, 4096), "c3": (8, 1, 1, 131072, 2048), "c2": (1, 1, 1, 65536, 1024)}[wl]
    ins, outs, groups, taps, B = cfg
    dev = torch.device("cuda", 0)
    eng = _Engine(np.float32, groups, ins, outs, 2 * B, taps, 0, 0, 0)
    eng.set_reset_offset(0)
    ir = torch.randn(taps, device=dev)
    for o in range(outs):
        for i in range(ins):
            eng.set_ir_device(0, i, o, ir.data_ptr(), taps)
    torch.cuda.synchronize()
    x = np.random.rand(ins, B).astype(np.float32)
    y = np.zeros((outs, B), np.float32)
    xr = [x[r] for r in range(ins)]
    yr = [y[r] for r in range(outs)]
    for mode in (1, 0, 1):
        eng.set_host_pipeline(bool(mode))
        for _ in range(3):
            eng.process(xr, yr, B)
        ts = []
        t_all = time.perf_counter()
        for k in range(20):
            t0 = time.perf_counter()
            eng.process(xr, yr, B)
            ts.append((time.perf_counter() - t0) * 1e3)
        t_all = (time.perf_counter() - t_all) * 1e3
        torch.cuda.synchronize()
        print("pipelined=%d  total %.2f ms for 20 calls; per call ms: %s" % (mode, t_all, " ".join("%.2f" % t for t in ts)))
    eng.close()

main()
