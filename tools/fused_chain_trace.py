"""Timeline of overlapping fused hops (hb_conv_set_hop_overlap): per hop, when the rank-0 CTA of the cluster enters, has its
products done, passes the cluster barriers, finishes the inverse transform and exits -- relative to the first hop shown.

    python tools/fused_chain_trace.py [c1|c2|c3] [mode 0|2] [calls]
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hisstools_library_b200.convolve import _Engine

CFG = {"c1": (1, 1, 1, 4096, 512), "c2": (1, 1, 1, 65536, 1024), "c3": (8, 1, 1, 131072, 2048)}


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
    mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    calls = int(sys.argv[3]) if len(sys.argv) > 3 else 400
    ins, outs, groups, taps, B = CFG[wl]
    dev = torch.device("cuda", 0)
    eng = _Engine(np.float32, groups, ins, outs, 2 * B, taps, 0, 0, 0)
    eng.set_reset_offset(0)
    eng.set_hop_overlap(mode)
    ir = torch.randn(taps, device=dev)
    for i in range(ins):
        eng.set_ir_device(0, i, 0, ir.data_ptr(), taps)
    x = torch.rand(4, ins, B, device=dev)
    y = torch.zeros(4, B, device=dev)
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    eng.set_trace(True)
    xp, yp, sp = [x[k].data_ptr() for k in range(4)], [y[k].data_ptr() for k in range(4)], st.cuda_stream
    for k in range(calls):
        eng.process_device(xp[k % 4], B, yp[k % 4], B, B, False, sp)
    tr, hop = eng.get_trace()
    print("workload %s, schedule %s, hop overlap mode %d, %d back-to-back calls, hops so far %d" % (wl, eng.schedule, mode, calls, hop))
    print("rank-0 CTA of the cluster (CTA 0), microseconds from the entry of the first hop shown; last = last CTA of the cluster")
    print("hop    entry    waited   fwd done   fwd+tail   products   barrier1   reduced   inverse    exit  | last-CTA entry  exit | entry-to-entry")
    t0, prev = None, None
    for h in range(hop - 12, hop):
        e = tr[h % 16]
        if e[0, 0, 0] == 0:
            continue
        if t0 is None:
            t0 = int(e[0, 0, 0])
        f = lambda v: (int(v) - t0) / 1e3 if v else float("nan")
        ncta = int((e[0, 0] > 0).sum())
        ent = f(e[0, 0, 0])
        print("%4d %8.2f %8.2f %9.2f %9.2f %10.2f %10.2f %9.2f %9.2f %8.2f | %8.2f %8.2f | %6.2f" % (
            h, ent, f(e[4, 0, 0]), f(e[4, 1, 0]), f(e[1, 0, 0]), f(e[1, 1, 0]), f(e[2, 0, 0]), f(e[2, 1, 0]), f(e[3, 0, 0]), f(e[0, 1, 0]),
            f(e[0, 0, ncta - 1]), f(e[0, 1, ncta - 1]), ent - prev if prev is not None else float("nan")))
        prev = ent


if __name__ == "__main__":
    main()
