"""Config 5 through the host-pointer call: microseconds per call by tail streams / host pipeline, the per-call wall-clock
pattern, and the device-side profile of the same calls (hb_conv_set_profiling) -- where does the period go?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from hisstools_library_b200.convolve import _Engine

groups, B, P = 16, 8192, 128
taps = B * P
dev = torch.device("cuda", 0)


def run(tail_streams, pipeline, profile, calls=300):
    eng = _Engine(np.float64, groups, 1, 1, 2 * B, taps, 0, 0, 0)
    eng.set_reset_offset(0)
    eng.set_tail_streams(tail_streams)
    ir = torch.randn(taps, device=dev, dtype=torch.float64) * 0.01
    for g in range(groups):
        eng.set_ir_device(g, 0, 0, ir.data_ptr(), taps)
    torch.cuda.synchronize()
    eng.set_host_pipeline(pipeline)
    x = [np.random.rand(groups, B) for _ in range(4)]
    y = np.zeros((groups, B))
    px = [eng.row_pointers([xi[r] for r in range(groups)]) for xi in x]
    py = eng.row_pointers([y[r] for r in range(groups)])
    for k in range(20):
        eng.process_pointers(px[k % 4], py, B)
    torch.cuda.synchronize()
    if profile:
        eng.set_profiling(True)
    ts = []
    t0 = time.perf_counter()
    for k in range(calls):
        a = time.perf_counter()
        eng.process_pointers(px[k % 4], py, B)
        ts.append((time.perf_counter() - a) * 1e6)
    torch.cuda.synchronize()
    per = (time.perf_counter() - t0) / calls * 1e6
    prof = eng.get_profile() if profile else None
    # the same engine, device-resident, for the period the GPU alone allows
    xd = torch.rand(groups, B, device=dev, dtype=torch.float64)
    yd = torch.zeros(groups, B, device=dev, dtype=torch.float64)
    st = torch.cuda.Stream()
    if profile:
        eng.set_profiling(False)
    for k in range(20):
        eng.process_device(xd.data_ptr(), B, yd.data_ptr(), B, B, False, st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for k in range(calls):
        eng.process_device(xd.data_ptr(), B, yd.data_ptr(), B, B, False, st.cuda_stream)
    eng.join(st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    dev_per = e0.elapsed_time(e1) / calls * 1e3
    print("tail streams %d  host pipeline %d  profiling %d:  %.1f us per host call (%.0f M samples/s), device-resident %.1f us" %
          (eng.tail_streams, pipeline, profile, per, groups * B / per, dev_per))
    print("   per-call wall clock, calls 100..115: " + " ".join("%.0f" % t for t in ts[100:116]))
    if prof:
        ms, hops = prof
        print("   device profile over %d hops (us per hop): %s" % (hops, {k: round(v / max(hops, 1) * 1e3, 1) for k, v in ms.items()}))
    eng.close()


if len(sys.argv) > 1 and sys.argv[1] == "trace":
    run(0, True, False, calls=2000)          # with HB_HOST_TRACE=1 in the environment: the host-side phases are printed at exit
else:
    run(0, True, False)
    run(1, True, False)
    run(0, True, True)
    run(0, False, False)
