"""Parity of the BENCHMARKED engines at BASELINE.json's stated sizes against the compiled reference (oracle/_ref):
the tile shapes, ring depths, eviction policies and schedules of the CUDA path depend on the size of the object
(plan_geometry), so the reduced shapes of test_gpu_conv.py do not walk every branch the benchmark runs.

  config 3  NToMonoConvolve 8 -> 1, 131072 taps, 2048-sample blocks, float  (whole object)
  config 4  Convolver 64 x 64, 262144 taps, 4096-sample blocks, float       (the full 8 GiB engine; 4 output rows x all
            64 inputs checked: about 2 s of reference CPU time)
  config 5  16 double channels, 1048576 taps, 8192-sample blocks            (3 channels checked, <= 1e-12)

Streams are P + 16 blocks long (delay line full); rel-RMS over the whole stream and over the last 16 blocks.
The synthetic data are bench.py's (same seeds), so these are the numbers bench.py prints as "parity".
"""
import numpy as np
import pytest

import checkers as ck

pytestmark = pytest.mark.gpu

TOL32, TOL64 = 1e-5, 1e-12


@pytest.fixture(scope="module")
def torch_dev():
    import torch
    if ck.ref() is None:
        pytest.skip("compiled reference not shipped")
    return torch, torch.device("cuda", 0)


def _load_matrix(torch, dev, eng, ins, outs, groups, taps, tdt, keep_rows):
    """bench.py's synthetic IRs on every pair; returns host copies of the rows (or banks) in keep_rows"""
    import bench
    gen = torch.Generator(device=dev)
    decay = torch.exp(-6.9 * torch.arange(taps, device=dev, dtype=torch.float64) / taps).to(tdt)
    kept = {}
    for g in range(groups):
        for o in range(outs):
            for i in range(ins):
                ir = bench.device_ir(gen, bench.ir_seed(ins, outs, g, o, i), taps, decay, tdt)
                assert eng.set_ir_device(g, i, o, ir.data_ptr(), taps) == 0
                key = g if groups > 1 else o
                if key in keep_rows:
                    kept[(key, i)] = ir.cpu().numpy()
    torch.cuda.synchronize()
    return gen, kept


def _stream_device(torch, eng, xs, rows_out, B, calls, check_rows, tdt):
    """xs [rows_in, n] on the device through process_device in calls of `calls` blocks; returns the checked rows"""
    n = xs.shape[1]
    y = torch.zeros(rows_out, n, device=xs.device, dtype=tdt)
    stream = torch.cuda.Stream(device=xs.device)
    with torch.cuda.stream(stream):
        pos, k = 0, 0
        while pos < n:
            m = min(calls[k % len(calls)] * B, n - pos)
            eng.process_device(xs[:, pos:].data_ptr(), xs.stride(0), y[:, pos:].data_ptr(), y.stride(0), m, False, stream.cuda_stream)
            pos += m
            k += 1
    torch.cuda.synchronize()
    return y[check_rows].cpu().numpy()


def _assert_rows(got, want, B, tol):
    for q in range(len(want)):
        assert ck.rel_rms(got[q], want[q]) <= tol, q
        assert ck.rel_rms(got[q][-16 * B:], want[q][-16 * B:]) <= tol, q


def test_config4_full_size_against_reference(torch_dev):
    """The benchmarked config-4 engine (64 x 64 pairs of 262144 taps: 8 GiB of spectra, overlapped schedule, TMA ring)
    on P + 16 = 80 blocks: one block per call (the timed path), calls of 4 and 8 blocks (multi-hop reuse), mixed call
    sizes, and the host-pointer path -- outputs 0, 21, 42, 63 against the reference's rows of 64 MonoConvolves."""
    torch, dev = torch_dev
    from hisstools_library_b200.convolve import _Engine
    import bench
    ins, outs, B, P = 64, 64, 4096, 64
    taps, hops = B * P, P + 16
    check = [0, 21, 42, 63]
    eng = _Engine(np.float32, 1, ins, outs, 2 * B, taps, 0, 0, 0)
    eng.set_reset_offset(0)
    gen, kept = _load_matrix(torch, dev, eng, ins, outs, 1, taps, torch.float32, set(check))
    assert eng.partitions == P
    pool = bench.input_pool(gen, 0, ins, B, 4, torch.float32, dev)
    xs = torch.cat([pool[k % 4] for k in range(hops)], dim=1).contiguous()
    irs = np.stack([np.stack([kept[(o, i)] for i in range(ins)]) for o in check])
    want = ck.ref_matrix_run(irs, xs.cpu().numpy(), 2 * B)
    got = {}
    for name, calls in (("one block per call", [1]), ("4 blocks per call", [4]), ("8 blocks per call", [8]), ("mixed", [1, 3, 8, 2, 1, 5])):
        eng.reset()
        got[name] = _stream_device(torch, eng, xs, outs, B, calls, check, torch.float32)
        _assert_rows(got[name], want, B, TOL32)
    assert eng.schedule == "overlapped"
    # the host-pointer (e2e) path: pipelined single-block calls, then a multi-block call
    eng.reset()
    xh = xs.cpu().numpy()
    yh = np.zeros((outs, hops * B), np.float32)
    for k in range(hops - 8):
        eng.process([xh[r, k * B:(k + 1) * B] for r in range(ins)], [yh[r, k * B:(k + 1) * B] for r in range(outs)], B)
    k0 = (hops - 8) * B
    eng.process([xh[r, k0:] for r in range(ins)], [yh[r, k0:] for r in range(outs)], 8 * B)
    _assert_rows(yh[check], want, B, TOL32)
    eng.close()


def test_config3_full_size_against_reference(torch_dev):
    """config 3 as the reference builds it: 8 MonoConvolve(131072, false, 4096) summed (NToMonoConvolve.cpp:35-43)."""
    torch, dev = torch_dev
    import hisstools_library_b200 as hb
    n_in, B, P = 8, 2048, 64
    L, hops = B * P, P + 16
    irs = np.stack([ck.synth_ir(L, 3000 + i) for i in range(n_in)])
    xs = np.stack([ck.synth_audio(B * hops, 3000 + i) for i in range(n_in)])
    want = ck.ref_matrix_run(irs[None], xs, 2 * B)[0]
    nm = hb.NToMonoConvolve(n_in, L, False, 2 * B)
    nm.setResetOffset(0)
    for i in range(n_in):
        assert nm.set(i, irs[i], L, False) == 0
    out = np.zeros(B * hops, np.float32)
    tmp = np.zeros(B, np.float32)
    for pos in range(0, B * hops, B):
        nm.process([x[pos:pos + B] for x in xs], out[pos:pos + B], tmp, B, n_in)
    assert ck.rel_rms(out, want) <= TOL32
    assert ck.rel_rms(out[-16 * B:], want[-16 * B:]) <= TOL32
    # the same object fed 4 blocks per call and ragged calls
    for sizes in ([4 * B], [B + 7, 3 * B - 7, 100, B - 100]):
        nm.reset()
        out2 = np.zeros(B * hops, np.float32)
        pos, k = 0, 0
        while pos < B * hops:
            m = min(sizes[k % len(sizes)], B * hops - pos)
            nm.process([x[pos:pos + m] for x in xs], out2[pos:pos + m], np.zeros(m, np.float32), m, n_in)
            pos += m
            k += 1
        assert ck.rel_rms(out2, want) <= TOL32


def test_config5_full_size_against_reference(torch_dev):
    """config 5: 16 independent double channels of 1048576 taps, 8192-sample blocks (transforms on clusters, 2-stage
    64 KiB ring); channels 0, 7, 15 against the restated double loop over the reference's own double FFT, <= 1e-12."""
    torch, dev = torch_dev
    from hisstools_library_b200.convolve import _Engine
    import bench
    groups, B, P = 16, 8192, 128
    taps, hops = B * P, P + 16
    check = [0, 7, 15]
    eng = _Engine(np.float64, groups, 1, 1, 2 * B, taps, 0, 0, 0)
    eng.set_reset_offset(0)
    gen, kept = _load_matrix(torch, dev, eng, 1, 1, groups, taps, torch.float64, set(check))
    assert eng.partitions == P
    pool = bench.input_pool(gen, 0, groups, B, 4, torch.float64, dev)
    xs = torch.cat([pool[k % 4] for k in range(hops)], dim=1).contiguous()
    xh = xs.cpu().numpy()
    want = np.stack([ck.ref_restated_run_f64(kept[(g, 0)], xh[g], 2 * B) for g in check])
    for calls in ([1], [4], [1, 2, 5]):
        eng.reset()
        got = _stream_device(torch, eng, xs, groups, B, calls, check, torch.float64)
        _assert_rows(got, want, B, TOL64)
    assert eng.fft_path == 2 and eng.schedule == "overlapped"
    eng.close()


def test_bench_parity_leg_runs_on_a_small_workload(torch_dev, capsys):
    """bench.py end to end on config 2 (seconds): the JSON line carries parity (ok), roofline, e2e, cpu_baseline, and
    the steady-state window holds as many tail launches as blocks (kernel_share_of_step <= 1)."""
    import json
    import subprocess
    import sys
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "c2", "--steps", "5", "--warmup", "3", "--cpu-seconds", "2",
                          "--min-seconds", "0.2"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["parity"]["ok"] and line["parity"]["rel_rms"] <= TOL32
    assert line["e2e"]["value"] > 0 and line["gpu_launches"] > 0 and line["roofline"]["achieved"] > 0
    assert line["cpu_baseline"]["kind"] == "reference"
    assert line["timing"]["timed_region_s"] >= 0.1            # (asked for 0.2 s; the calibration run is a little slower than the steady state)
