#!/usr/bin/env python
"""The public FFT family timed the way the reference's own tester does it (`- Test/FFT_Tester/FFT_Tester/main.cpp:142-199`:
many transforms of one size, fft / ifft / rfft / rifft, float), next to the compiled reference on one host core.

  host call   hisstools_fft / ifft / rfft / rifft(setup, split, log2n) through the C ABI: one H2D copy, one kernel, one D2H copy per
              transform -- the drop-in path, bound by the round trip, not by the transform
  batched     hb_rfft_real_batched_dev / hb_rifft_real_batched_dev on device-resident buffers, 4096 transforms per launch -- what
              the transform kernels themselves deliver
Prints a table (transforms per second).  Run on the GPU box: python tools/fft_family_bench.py > gpurun_out/fft_family.txt"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                                                    # noqa: E402  (device buffers and events only)
import checkers as ck                                           # noqa: E402
import hisstools_library_b200 as hb                             # noqa: E402
from hisstools_library_b200 import _abi                         # noqa: E402


def main():
    ref = ck.ref()
    lib = _abi.lib()
    reps = 2000
    print("# float transforms per second; reference = compiled reference, one host core (SSE2 -O2); %d transforms per figure" % reps)
    print("%-7s %-6s %14s %14s %16s" % ("log2n", "kind", "reference CPU", "host call", "batched (device)"))
    for log2n in (8, 10, 12, 13, 14, 16):
        n = 1 << log2n
        setup = hb.hisstools_create_setup(log2n)
        rs = ref.ref_fft_setup_f32(log2n) if ref is not None else None
        rng = np.random.default_rng(log2n)
        for kind in ("fft", "ifft", "rfft", "rifft"):
            pts = n if kind in ("fft", "ifft") else n // 2
            re, im = rng.uniform(-1, 1, pts).astype(np.float32), rng.uniform(-1, 1, pts).astype(np.float32)
            # reference
            r_rate = float("nan")
            if ref is not None:
                fn = getattr(ref, "ref_%s_f32" % kind)
                a, b = re.copy(), im.copy()
                t0 = time.perf_counter()
                for _ in range(reps):
                    fn(rs, ck.fptr(a), ck.fptr(b), log2n)
                    a[:] = re; b[:] = im                     # keep the values bounded (the tester refills too)
                r_rate = reps / (time.perf_counter() - t0)
            # ours, host call (same in-place signature)
            split = hb.Split(re.copy(), im.copy())
            f = getattr(hb, "hisstools_" + kind)
            for _ in range(20):
                f(setup, split, log2n)
            t0 = time.perf_counter()
            for _ in range(reps):
                f(setup, split, log2n)
            h_rate = reps / (time.perf_counter() - t0)
            # ours, batched on the device (real transforms only: the entry points the convolver's kernels share)
            b_rate = float("nan")
            if kind in ("rfft", "rifft") and log2n <= 15:
                batch = 4096
                x = torch.rand(batch, n, device="cuda") * 2 - 1
                pr, pi = torch.zeros(batch, n // 2, device="cuda"), torch.zeros(batch, n // 2, device="cuda")
                stream = torch.cuda.Stream()                 # a real stream: handle 0 would mean the setup's own stream
                torch.cuda.set_stream(stream)
                st = stream.cuda_stream
                h = setup._h if hasattr(setup, "_h") else setup.handle
                def run():
                    if kind == "rfft":
                        return lib.hb_rfft_real_batched_dev(h, C.c_void_p(x.data_ptr()), C.c_void_p(pr.data_ptr()), C.c_void_p(pi.data_ptr()), log2n, batch, n, n // 2, C.c_void_p(st))
                    return lib.hb_rifft_real_batched_dev(h, C.c_void_p(pr.data_ptr()), C.c_void_p(pi.data_ptr()), C.c_void_p(x.data_ptr()), log2n, batch, n // 2, n, C.c_void_p(st))
                assert run() == 0, _abi.last_error()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(10):
                    assert run() == 0, _abi.last_error()
                e1.record(stream)
                torch.cuda.synchronize()
                b_rate = 10 * batch / (e0.elapsed_time(e1) * 1e-3)
            print("%-7d %-6s %14.0f %14.0f %16.0f" % (log2n, kind, r_rate, h_rate, b_rate))
        if ref is not None:
            ref.ref_fft_setup_free_f32(rs)
        hb.hisstools_destroy_setup(setup)


if __name__ == "__main__":
    main()
