"""Kernel timeline of a few hops at a bench workload (hb_conv_set_trace): when do the CTAs of the forward FFT,
head / tail multiply-accumulate and inverse FFT kernels really start and end?

    python tools/trace_timeline.py [c4|c3|c5] [overlapped|serial]
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hisstools_library_b200.convolve import _Engine

CFG = {"c1": (1, 1, 1, 4096, 512, np.float32), "c2": (1, 1, 1, 65536, 1024, np.float32),
       "c4": (64, 64, 1, 262144, 4096, np.float32), "c4n8": (8, 64, 1, 262144, 4096, np.float32), "c3": (8, 1, 1, 131072, 2048, np.float32),
       "c5": (1, 1, 16, 1048576, 8192, np.float64)}
KIND = ["fwd", "head", "tail", "inv", "gath"]


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
    sched = sys.argv[2] if len(sys.argv) > 2 else "overlapped"
    sync_each = len(sys.argv) > 3 and sys.argv[3] == "sync"        # idle GPU before every hop: a late inverse FFT beside a resident tail
    variant = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    ins, outs, groups, taps, B, dt = CFG[wl]
    dev = torch.device("cuda", 0)
    tdt = torch.float64 if dt == np.float64 else torch.float32
    eng = _Engine(dt, groups, ins, outs, 2 * B, taps, 0, 0, 0)
    eng.set_reset_offset(0)
    eng.set_schedule(None if sched == "auto" else sched == "overlapped")
    eng.set_tuning(0, variant)
    ir = torch.randn(taps, device=dev, dtype=tdt)
    for g in range(groups):
        for o in range(outs):
            for i in range(ins):
                eng.set_ir_device(g, i, o, ir.data_ptr(), taps)
    x = torch.rand(groups * ins, B, device=dev, dtype=tdt)
    y = torch.zeros(groups * outs, B, device=dev, dtype=tdt)
    st = torch.cuda.Stream()
    for _ in range(5):
        eng.process_device(x.data_ptr(), B, y.data_ptr(), B, B, False, st.cuda_stream)
    torch.cuda.synchronize()
    eng.set_trace(True)
    for _ in range(12):
        eng.process_device(x.data_ptr(), B, y.data_ptr(), B, B, False, st.cuda_stream)
        if sync_each:
            torch.cuda.synchronize()
    tr, hop = eng.get_trace()
    print("workload %s, schedule %s, variant %d, %s, hops so far %d" % (wl, eng.schedule, variant, "device idle before every hop" if sync_each else "back to back", hop))
    hops = [h for h in range(hop - 8, hop - 1)]
    t0 = None
    for h in hops:
        for k in range(5):
            ent, ext = tr[h % 16, k, 0], tr[h % 16, k, 1]
            m = ent > 0
            if not m.any():
                continue
            if t0 is None:
                t0 = int(ent[m].min())
            e, x_ = ent[m].astype(np.int64) - t0, ext[m].astype(np.int64) - t0
            print("hop %3d %-4s ctas %3d  first entry %9.1f us  last entry %9.1f us  first exit %9.1f us  last exit %9.1f us" %
                  (h, KIND[k], int(m.sum()), e.min() / 1e3, e.max() / 1e3, x_[x_ > -t0 // 2].min() / 1e3 if (x_ > -t0 // 2).any() else -1, x_.max() / 1e3))
    if os.environ.get("HB_TRACE_DETAIL"):
        # per-CTA entry / exit of one kind in the last complete hop (e.g. HB_TRACE_DETAIL=inv)
        k = KIND.index(os.environ["HB_TRACE_DETAIL"])
        h = hop - 2
        ent, ext = tr[h % 16, k, 0].astype(np.int64), tr[h % 16, k, 1].astype(np.int64)
        idx = np.nonzero(ent > 0)[0]
        print("hop %d %s, per CTA (entry, exit) in us:" % (h, KIND[k]))
        print(" ".join("%d:(%.1f,%.1f)" % (i, (ent[i] - t0) / 1e3, (ext[i] - t0) / 1e3) for i in idx))
    eng.close()


main()
