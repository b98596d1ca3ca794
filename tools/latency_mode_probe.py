"""Throughput of the reference's shipped partition schemes (Convolver(numIns, numOuts, LatencyMode)) on the GPU next to the uniform
scheme of BASELINE config 4: python tools/latency_mode_probe.py [ins] [outs] [taps] [block]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hisstools_library_b200.convolve import _Matrix, _scheme


def main():
    ins = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    outs = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    taps = int(sys.argv[3]) if len(sys.argv) > 3 else 262144
    n = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(1)
    ir = (rng.standard_normal(taps) * np.exp(-6.9 * np.arange(taps) / taps)).astype(np.float32)
    x = torch.rand(ins, n, device=dev) * 2 - 1
    y = torch.zeros(outs, n, device=dev)
    st = torch.cuda.Stream()
    for name, scheme in (("uniform FFT 8192", (False, 8192)), ("kLatencyZero", (True, 256, 1024, 4096, 16384)), ("kLatencyShort", (False, 256, 1024, 4096, 16384)),
                         ("kLatencyMedium", (False, 1024, 4096, 16384))):
        m = _Matrix(1, ins, outs, taps, _scheme(scheme), np.float32, 0)
        m.setResetOffset(0)
        t0 = time.perf_counter()
        for o in range(outs):
            for i in range(ins):
                m.set(0, i, o, ir, taps, True)
        torch.cuda.synchronize()
        t_set = time.perf_counter() - t0
        for _ in range(3):
            m.process_device(x.data_ptr(), n, y.data_ptr(), n, n, False, st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 20
        e0.record(st)
        for _ in range(steps):
            m.process_device(x.data_ptr(), n, y.data_ptr(), n, n, False, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        # end to end: host rows in, host rows out (hb_matrix_process: what Convolver::process costs a caller), wall clock
        xh = [np.ascontiguousarray(r) for r in x.cpu().numpy()]
        yh = [np.zeros(n, np.float32) for _ in range(outs)]
        for _ in range(3):
            m.process(xh, yh, n, False)
        t0 = time.perf_counter()
        for _ in range(steps):
            m.process(xh, yh, n, False)
        ms_host = (time.perf_counter() - t0) / steps * 1e3
        print("%-18s end to end from host rows: %.3f ms per block = %.1f M output-samples/s" % (name, ms_host, outs * n / ms_host / 1e3))
        print("%-18s %dx%d, %d taps, blocks of %d: %.3f ms per block = %.1f M output-samples/s  (IR load %.1f s; parts %s, head %d taps; schedules %s)" %
              (name, ins, outs, taps, n, ms, outs * n / ms / 1e3, t_set, [e.fft_size for e in m.engines], m.head_taps, [e.schedule for e in m.engines]))
        m.close()


if not (len(sys.argv) > 1 and sys.argv[1] == "parts"):
    main()


def per_part():
    """time every part of a kLatencyShort matrix on its own (borrowed engines)"""
    ins, outs, taps, n = 64, 64, 262144, 4096
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(1)
    ir = (rng.standard_normal(taps) * np.exp(-6.9 * np.arange(taps) / taps)).astype(np.float32)
    x = torch.rand(ins, n, device=dev) * 2 - 1
    y = torch.zeros(outs, n, device=dev)
    st = torch.cuda.Stream()
    m = _Matrix(1, ins, outs, taps, _scheme((False, 256, 1024, 4096, 16384)), np.float32, 0)
    m.setResetOffset(0)
    for o in range(outs):
        for i in range(ins):
            m.set(0, i, o, ir, taps, True)
    torch.cuda.synchronize()
    for e in m.engines:
        for mh in (True, False):
            e.set_multi_hop(mh)
            for _ in range(4):
                e.process_device(x.data_ptr(), n, y.data_ptr(), n, n, False, st.cuda_stream)
            torch.cuda.synchronize()
            lib_launch0 = __import__("hisstools_library_b200")._abi.lib().hb_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(20):
                e.process_device(x.data_ptr(), n, y.data_ptr(), n, n, False, st.cuda_stream)
            e1.record(st)
            torch.cuda.synchronize()
            launches = (__import__("hisstools_library_b200")._abi.lib().hb_launch_count() - lib_launch0) / 20
            print("part FFT %5d  partitions %3d  multi_hop %-5s schedule %-10s %.3f ms per block of %d, %.0f launches per block" %
                  (e.fft_size, e.partitions, mh, e.schedule, e0.elapsed_time(e1) / 20, n, launches))
    m.close()


if len(sys.argv) > 1 and sys.argv[1] == "parts":
    per_part()
