"""GPU parity of the partitioned-convolution path (CUDA through the C ABI) against
 * the golden vectors produced by the unmodified reference (tests/golden/),
 * the unmodified reference compiled in place (oracle/_ref, shipped prebuilt to the GPU box),
 * the plain-C oracle (oracle/hiss_oracle.c) where the reference has no class (double engine),
and through size-independent properties at BASELINE.json's full sizes.
Tolerances (north_star): <= 1e-5 relative RMS float, <= 1e-12 double.
"""
import os

import numpy as np
import pytest

import checkers as ck

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz"))
TOL32, TOL64 = 1e-5, 1e-12


@pytest.fixture(scope="module")
def hb():
    import hisstools_library_b200 as h
    return h


def stream(obj_process, x, block, dtype=np.float32, sizes=None):
    """run x through process(in, out, n) in `block`-sample calls (or the given call sizes)."""
    y = np.zeros(len(x), dtype)
    pos, k = 0, 0
    while pos < len(x):
        n = min(sizes[k % len(sizes)] if sizes else block, len(x) - pos)
        obj_process(x[pos:pos + n], y[pos:pos + n], n)
        pos += n
        k += 1
    return y


# ---- PartitionedConvolve ------------------------------------------------------------------------

@pytest.mark.parametrize("schedule", ["overlapped", "serial", "auto"])
@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("name", ["c1", "ragged", "phase", "slice", "trunc", "min"])
def test_pconv_golden(hb, name, variant, schedule):
    fft, block, max_len, offset, length, reset_offset, err = (int(v) for v in G["pconv_%s_meta" % name])
    ir, x, want = G["pconv_%s_ir" % name], G["pconv_%s_x" % name], G["pconv_%s_y" % name]
    pc = hb.PartitionedConvolve(fft, len(ir) if max_len < 0 else max_len, offset, length)
    pc.engine.set_tuning(0, variant)
    pc.engine.set_schedule(None if schedule == "auto" else schedule == "overlapped")
    pc.setResetOffset(reset_offset)
    assert int(pc.set(ir, len(ir))) == err
    got = stream(lambda a, b, n: pc.process(a, b, n), x, block)
    assert ck.rel_rms(got, want) <= TOL32
    if schedule == "auto":
        assert pc.engine.schedule == "fused"             # a single small channel: one cluster launch per hop


@pytest.mark.parametrize("variant", [1, 0])
def test_pconv_config2_against_reference(hb, variant):
    """BASELINE config 2: 1ch, 65536 taps, 1024-sample blocks = PartitionedConvolve(2048, 65536, 0, 0)."""
    if ck.ref() is None:
        pytest.skip("compiled reference not shipped")
    ir = ck.synth_ir(65536, 0)
    x = ck.synth_audio(1024 * (64 + 32), 0)
    want, _ = ck.ref_pconv_run(2048, ir, x, 1024)
    pc = hb.PartitionedConvolve(2048, 65536, 0, 0)
    pc.engine.set_tuning(0, variant)
    pc.setResetOffset(0)
    assert int(pc.set(ir, len(ir))) == 0
    got = stream(lambda a, b, n: pc.process(a, b, n), x, 1024)
    assert ck.rel_rms(got, want) <= TOL32
    assert ck.rel_rms(got[-16 * 1024:], want[-16 * 1024:]) <= TOL32          # FDL full
    truth = ck.direct_convolve_delayed(ir[:4096], x[:8192], 1024)            # delay is exactly B


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_schedules_agree_and_survive_mid_stream_changes(hb, dtype):
    """Overlapped schedule (tail of the next hop computed ahead on a second stream) against the serial one on a
    4-in x 3-out matrix: same result up to summation order, also across reset(), a new IR mid-stream (the tail
    launched ahead must be discarded) and a ragged call pattern."""
    from hisstools_library_b200.convolve import _Engine
    n_in, n_out, L, B = 4, 3, 2500, 256
    tol = TOL32 if dtype == np.float32 else TOL64
    irs = [[ck.synth_ir(L, 500 + 10 * o + i).astype(dtype) for i in range(n_in)] for o in range(n_out)]
    irs2 = [[ck.synth_ir(L // 2, 900 + 10 * o + i).astype(dtype) for i in range(n_in)] for o in range(n_out)]
    xs = np.stack([ck.synth_audio(B * 40 + 77, 500 + i) for i in range(n_in)]).astype(dtype)
    outs = {}
    for schedule in ("overlapped", "serial"):
        e = _Engine(dtype, 1, n_in, n_out, 2 * B, L, 0, 0, 0)
        e.set_schedule(schedule == "overlapped")
        e.set_reset_offset(0)
        for o in range(n_out):
            for i in range(n_in):
                e.set_ir(0, i, o, irs[o][i], L)
        y = np.zeros((n_out, xs.shape[1]), dtype)
        pos, k = 0, 0
        sizes = [B, B, 100, 3 * B, 1, 2 * B + 5, B]
        while pos < xs.shape[1]:
            n = min(sizes[k % len(sizes)], xs.shape[1] - pos)
            if k == 9:
                e.reset()
            if k == 17:
                for o in range(n_out):
                    for i in range(n_in):
                        e.set_ir(0, i, o, irs2[o][i], L // 2)
            xi = [np.ascontiguousarray(xs[r, pos:pos + n]) for r in range(n_in)]
            yo = [np.zeros(n, dtype) for _ in range(n_out)]
            e.process(xi, yo, n)
            for r in range(n_out):
                y[r, pos:pos + n] = yo[r]
            pos += n
            k += 1
        assert e.schedule == schedule
        outs[schedule] = y
        e.close()
    for o in range(n_out):
        assert ck.rel_rms(outs["overlapped"][o], outs["serial"][o]) <= (1e-6 if dtype == np.float32 else 1e-14)
    # and the first stretch (before the reset) against the truth
    first = sum([B, B, 100, 3 * B, 1, 2 * B + 5, B, B, B])
    for o in range(n_out):
        truth = sum(ck.direct_convolve_delayed(irs[o][i], xs[i][:first], B) for i in range(n_in))
        assert ck.rel_rms(outs["overlapped"][o][:first], truth) <= tol * (1 if dtype == np.float32 else 10)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("ins,outs,groups,B,L,fused", [(1, 1, 1, 512, 4096, True), (1, 1, 1, 1024, 65536, True), (8, 1, 1, 2048, 30000, True),
                                                       (8, 1, 1, 2048, 131072, True), (8, 1, 1, 2048, 400000, False),
                                                       (3, 1, 2, 256, 5000, True), (1, 1, 3, 16, 100, True), (5, 1, 1, 64, 64, True),
                                                       (2, 2, 1, 512, 3000, True), (3, 5, 1, 256, 2000, True), (2, 8, 2, 128, 1000, True)])
def test_fused_hop_against_serial_and_truth(hb, dtype, ins, outs, groups, B, L, fused):
    """The fused single-launch hop (one thread-block cluster per output: hb_conv_fused.cuh) on small engines -- BASELINE
    configs 1, 2 and 3 (the last on a cluster of 16), several groups, tiny and one-partition sizes, stereo and wider matrices
    (every output's cluster transforms the inputs for itself) -- against the serial three-kernel hop (summation order only) and
    float64 direct convolution, ragged calls included; an 8 -> 1 engine with 51 MB of spectra per output is not eligible."""
    from hisstools_library_b200.convolve import _Engine
    if dtype == np.float64 and B > 2048:
        pytest.skip("spectrum above one bin tile")
    tol = TOL32 if dtype == np.float32 else TOL64
    if dtype == np.float64 and ins * L > 500000:
        fused = False                                        # twice the bytes per rank: past the limit of a cluster of 16
    irs = [[[ck.synth_ir(L, 700 + 100 * g + 10 * o + i).astype(dtype) for i in range(ins)] for o in range(outs)] for g in range(groups)]
    n = B * 12 + 37
    xs = np.stack([ck.synth_audio(n, 700 + r) for r in range(groups * ins)]).astype(dtype)
    res = {}
    for mode in ("auto", "serial"):
        e = _Engine(dtype, groups, ins, outs, 2 * B, L, 0, 0, 0)
        e.set_schedule(None if mode == "auto" else False)
        e.set_reset_offset(0)
        for g in range(groups):
            for o in range(outs):
                for i in range(ins):
                    e.set_ir(g, i, o, irs[g][o][i], L)
        y = np.zeros((groups * outs, n), dtype)
        pos, k = 0, 0
        sizes = [B, B, B // 2 + 1, 3 * B, 5, B]
        while pos < n:
            m = min(sizes[k % len(sizes)], n - pos)
            yo = [np.zeros(m, dtype) for _ in range(groups * outs)]
            e.process([np.ascontiguousarray(xs[r, pos:pos + m]) for r in range(groups * ins)], yo, m)
            for r in range(groups * outs):
                y[r, pos:pos + m] = yo[r]
            pos += m
            k += 1
        if mode == "serial":
            assert e.schedule == "serial"
        elif fused:
            assert e.schedule == "fused"
        else:
            assert e.schedule in ("overlapped", "serial")
        res[mode] = y
        e.close()
    for g in range(groups):
        for o in range(outs):
            r = g * outs + o
            assert ck.rel_rms(res["auto"][r], res["serial"][r]) <= tol / 5
            truth = sum(ck.direct_convolve_delayed_fft(irs[g][o][i], xs[g * ins + i], B) for i in range(ins))
            assert ck.rel_rms(res["auto"][r], truth) <= tol * (1 if dtype == np.float32 else 10)


@pytest.mark.parametrize("dtype,schedule", [(np.float32, None), (np.float32, False), (np.float64, None)])
def test_multi_hop_reuse_matches_hop_by_hop(hb, dtype, schedule):
    """Calls that bring several hops at once run batches of 4 / 2 hops over one pass of the IR spectra (k_cmac_tma_mh) on
    HBM-bound engines: a 4-in x 16-out matrix with 32 partitions, call sizes of 7, 3, 1, 4, 2 ... hops mixed with ragged
    calls, against the same engine with the batching off (summation order only) and float64 direct convolution; with the
    overlapped schedule around the batches (tail launched ahead after a batch, discarded before one) and the serial one."""
    from hisstools_library_b200.convolve import _Engine
    n_in, n_out, B = 4, 16, 256
    L = 32 * B
    tol = TOL32 if dtype == np.float32 else TOL64
    rng = np.random.default_rng(99)
    irs = (rng.standard_normal((n_out, n_in, L)) * np.exp(-6.9 * np.arange(L) / L)).astype(dtype)
    calls = [7 * B, B, 3 * B, 100, B - 100, 4 * B, 2 * B, B, 9 * B, 5, 6 * B + 5 - 10, 5, 8 * B]
    n = sum(calls)
    xs = np.stack([ck.synth_audio(n, 800 + i) for i in range(n_in)]).astype(dtype)
    outs = {}
    for mh in (True, False):
        e = _Engine(dtype, 1, n_in, n_out, 2 * B, L, 0, 0, 0)
        e.set_schedule(schedule)
        e.set_multi_hop(mh)
        e.set_reset_offset(0)
        for o in range(n_out):
            for i in range(n_in):
                e.set_ir(0, i, o, irs[o, i], L)
        y = np.zeros((n_out, n), dtype)
        pos = 0
        for m in calls:
            yo = [np.zeros(m, dtype) for _ in range(n_out)]
            e.process([np.ascontiguousarray(xs[r, pos:pos + m]) for r in range(n_in)], yo, m)
            for o in range(n_out):
                y[o, pos:pos + m] = yo[o]
            pos += m
        assert e.schedule == ("overlapped" if schedule is None else "serial")
        outs[mh] = y
        e.close()
    for o in range(n_out):
        assert ck.rel_rms(outs[True][o], outs[False][o]) <= tol / 5
    for o in (0, 7, 15):
        truth = sum(ck.direct_convolve_delayed_fft(irs[o, i], xs[i], B) for i in range(n_in))
        assert ck.rel_rms(outs[True][o], truth) <= tol * (1 if dtype == np.float32 else 10)


def test_pconv_call_sizes_do_not_matter(hb):
    """any chunking of the stream gives the same samples (SURVEY B: bit-identical in the reference)."""
    ir = ck.synth_ir(3000, 1)
    x = ck.synth_audio(256 * 40 + 5, 1)
    outs = []
    for sizes in ([256], [1], [7, 255, 256, 257, 768, 3], [100000]):
        if sizes == [1]:
            xs = x[:600]
        else:
            xs = x
        pc = hb.PartitionedConvolve(512, 3000, 0, 0)
        pc.setResetOffset(0)
        pc.set(ir)
        outs.append(stream(lambda a, b, n: pc.process(a, b, n), xs, 0, sizes=sizes))
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[0], outs[3])
    assert np.array_equal(outs[0][:600], outs[1])
    truth = ck.direct_convolve_delayed(ir, x, 256)
    assert ck.rel_rms(outs[0], truth) <= TOL32


@pytest.mark.parametrize("fft", [32, 64, 128, 256, 1024, 4096, 16384, 32768])
def test_pconv_fft_sizes(hb, fft):
    B = fft // 2
    L = min(5 * B + 3, 40000)
    ir = ck.synth_ir(L, 2)
    x = ck.synth_audio(B * 9 + 11, 2)
    pc = hb.PartitionedConvolve(fft, L, 0, 0)
    pc.setResetOffset(0)
    assert int(pc.set(ir)) == 0
    got = stream(lambda a, b, n: pc.process(a, b, n), x, max(B // 2, 1) + 1)
    if ck.ref() is not None:
        want, _ = ck.ref_pconv_run(fft, ir, x, B)
        assert ck.rel_rms(got, want) <= TOL32
    assert ck.rel_rms(got, ck.direct_convolve_delayed(ir, x, B)) <= TOL32


@pytest.mark.parametrize("fft,dtype", [(65536, np.float32), (131072, np.float32), (1 << 20, np.float32), (32768, np.float64), (65536, np.float64)])
def test_pconv_fft_sizes_above_one_cta(hb, fft, dtype):
    """FFT sizes whose half-length transform does not fit one CTA's shared memory (four-step path, hb_conv_big.cuh), up to
    the reference's largest (2^20, PartitionedConvolve.h:18-19): against the unmodified reference (float) / the C oracle
    (double) and float64 direct convolution, streamed in ragged calls; both schedules."""
    B = fft // 2
    L = 2 * B + B // 3 + 7                                   # three partitions, the last one ragged
    ir = ck.synth_ir(L, 6).astype(dtype)
    x = ck.synth_audio(B * 5 + 1234, 6).astype(dtype)
    tol = TOL32 if dtype == np.float32 else TOL64
    truth = ck.direct_convolve_delayed_fft(ir, x, B)
    outs = {}
    for schedule in (True, False):
        pc = hb.PartitionedConvolve(fft, L, 0, 0, dtype=dtype)
        pc.engine.set_schedule(schedule)
        pc.setResetOffset(0)
        assert int(pc.set(ir)) == 0
        got = stream(lambda a, b, n: pc.process(a, b, n), x, B // 2 + 77, dtype=dtype)
        assert ck.rel_rms(got, truth) <= tol
        outs[schedule] = got
        del pc
    assert ck.rel_rms(outs[True], outs[False]) <= tol / 10
    if dtype == np.float32 and ck.ref() is not None and fft <= 131072:
        want, _ = ck.ref_pconv_run(fft, ir, x, B)
        assert ck.rel_rms(outs[True], want) <= TOL32
    if dtype == np.float64 and fft <= 32768:
        want, _ = ck.oracle_pconv_run(fft, ir, x, B, dtype=np.float64)
        assert ck.rel_rms(outs[True], want) <= TOL64


def test_pconv_semantics(hb):
    E = hb.ConvolveError
    pc = hb.PartitionedConvolve(512, 1000, 0, 0)         # max length rounds up to 1024 (cpp:77-82)
    x = ck.synth_audio(2048, 3)
    y = np.full(2048, 7.0, np.float32)
    assert pc.process(x, y, 2048) is False and np.all(y == 7.0)               # no IR: untouched (cpp:262-263)
    assert pc.set(ck.synth_ir(1024, 3)) == E.CONVOLVE_ERR_NONE
    assert pc.set(ck.synth_ir(1281, 3)) == E.CONVOLVE_ERR_MEM_ALLOC_TOO_SMALL
    assert pc.setFFTSize(16) == E.CONVOLVE_ERR_FFT_SIZE_OUT_OF_RANGE
    assert pc.setFFTSize(1024) == E.CONVOLVE_ERR_FFT_SIZE_OUT_OF_RANGE
    assert pc.setFFTSize(300) == E.CONVOLVE_ERR_FFT_SIZE_NON_POWER_OF_TWO     # rounds up to 512: unchanged size
    assert pc.process(x, y, 512) is True
    assert pc.setFFTSize(256) == E.CONVOLVE_ERR_NONE                          # new size drops the IR (cpp:145-149)
    y[:] = 7.0
    assert pc.process(x, y, 512) is False and np.all(y == 7.0)
    assert pc.setLength(5000) == E.CONVOLVE_ERR_PARTITION_LENGTH_TOO_LARGE
    assert pc.setLength(0) == E.CONVOLVE_ERR_NONE
    # reset mid-stream restarts from silence
    ir = ck.synth_ir(700, 4)
    pc.setResetOffset(0)
    pc.set(ir)
    a = stream(lambda i, o, n: pc.process(i, o, n), x, 128)
    pc.reset()
    b = stream(lambda i, o, n: pc.process(i, o, n), x, 128)
    assert np.array_equal(a, b)
    assert ck.rel_rms(a, ck.direct_convolve_delayed(ir, x, 128)) <= TOL32


@pytest.mark.parametrize("schedule", ["overlapped", "serial"])
@pytest.mark.parametrize("dtype,B", [(np.float32, 2048), (np.float32, 8192), (np.float32, 16384), (np.float64, 2048), (np.float64, 4096), (np.float64, 8192)])
def test_fft_paths_agree(hb, dtype, B, schedule):
    """hb_conv_set_fft_path: transforms on clusters of 8 CTAs (distributed shared memory, hb_conv_cluster.cuh) and the
    four-step chains against one CTA per transform, on a 2-group 2-in x 3-out matrix with ragged call sizes (hand-over of
    the previous block, accumulation into the caller's rows), and against float64 direct convolution."""
    from hisstools_library_b200.convolve import _Engine
    groups, n_in, n_out = 2, 2, 3
    L = 3 * B + 100
    tol = TOL32 if dtype == np.float32 else TOL64
    irs = {(g, o, i): ck.synth_ir(L, 700 + 100 * g + 10 * o + i).astype(dtype) for g in range(groups) for o in range(n_out) for i in range(n_in)}
    xs = np.stack([ck.synth_audio(B * 6 + 77, 700 + i) for i in range(groups * n_in)]).astype(dtype)
    sizes = [B, B, 100, 2 * B, 1, B + 5, B]
    outs = {}
    for path in (1, 2, 3):
        if path == 3 and B < 4096:
            continue
        e = _Engine(dtype, groups, n_in, n_out, 2 * B, L, 0, 0, 0)
        e.set_schedule(schedule == "overlapped")
        e.set_multi_hop(False)
        e.set_fft_path(path)
        e.set_reset_offset(0)
        for (g, o, i), ir in irs.items():
            e.set_ir(g, i, o, ir, L)
        y = np.zeros((groups * n_out, xs.shape[1]), dtype)
        pos, k = 0, 0
        while pos < xs.shape[1]:
            n = min(sizes[k % len(sizes)], xs.shape[1] - pos)
            xi = [np.ascontiguousarray(xs[r, pos:pos + n]) for r in range(groups * n_in)]
            yo = [np.zeros(n, dtype) for _ in range(groups * n_out)]
            e.process(xi, yo, n)
            for r in range(groups * n_out):
                y[r, pos:pos + n] = yo[r]
            pos += n
            k += 1
        assert e.fft_path == path
        outs[path] = y
        e.close()
    for path in outs:
        for r in range(groups * n_out):
            assert ck.rel_rms(outs[path][r], outs[1][r]) <= (2e-6 if dtype == np.float32 else 1e-14), (path, r)
    for g in range(groups):
        for o in range(n_out):
            truth = sum(ck.direct_convolve_delayed_fft(irs[(g, o, i)], xs[g * n_in + i], B) for i in range(n_in))
            assert ck.rel_rms(outs[2][g * n_out + o], truth) <= tol * (1 if dtype == np.float32 else 10)


def test_pconv_double_engine(hb):
    """double engine against the restated double loop on the reference's own double FFT (golden) and
    against float64 direct convolution; BASELINE config 5 tolerance 1e-12."""
    ir, x, want = G["pconv64_ir"], G["pconv64_x"], G["pconv64_y"]
    pc = hb.PartitionedConvolve(512, len(ir), 0, 0, dtype=np.float64)
    pc.setResetOffset(0)
    assert int(pc.set(ir)) == 0
    got = stream(lambda a, b, n: pc.process(a, b, n), x, 256, dtype=np.float64)
    assert ck.rel_rms(got, want) <= TOL64
    assert ck.rel_rms(got, ck.direct_convolve_delayed(ir, x, 256)) <= TOL64


def test_pconv_double_config5_shape_one_channel(hb):
    """config-5 geometry (FFT 16384, B = 8192) on one channel with a shortened IR, vs the C oracle."""
    B = 8192
    ir = ck.synth_ir(5 * B + 100, 5).astype(np.float64)
    x = ck.synth_audio(B * 8, 5).astype(np.float64)
    want, _ = ck.oracle_pconv_run(2 * B, ir, x, B, dtype=np.float64)
    pc = hb.PartitionedConvolve(2 * B, len(ir), 0, 0, dtype=np.float64)
    pc.setResetOffset(0)
    pc.set(ir)
    got = stream(lambda a, b, n: pc.process(a, b, n), x, B, dtype=np.float64)
    assert ck.rel_rms(got, want) <= TOL64


# ---- MonoConvolve ---------------------------------------------------------------------------------

def test_mono_config1_against_reference(hb):
    """BASELINE config 1: MonoConvolve(4096, false, 1024), 512-sample blocks."""
    ir = ck.synth_ir(4096, 6)
    x = ck.synth_audio(512 * 48, 6)
    mc = hb.MonoConvolve(4096, False, 1024)
    mc.setResetOffset(0)
    assert mc.set(ir, len(ir), False) == 0
    tmp = np.zeros(512, np.float32)
    got = stream(lambda a, b, n: mc.process(a, tmp, b, n), x, 512)
    rl = ck.ref()
    if rl is not None:
        h = rl.ref_mono_create_custom(4096, 0, 1024, 0, 0, 0)
        rl.ref_mono_set_reset_offset(h, 0)
        rl.ref_mono_set(h, ck.fptr(ir), len(ir), 0)
        rl.ref_mono_set_reset_offset(h, 0)
        want, t = np.zeros_like(x), np.zeros_like(x)
        for pos in range(0, len(x), 512):
            rl.ref_mono_process(h, ck.fptr(x[pos:]), ck.fptr(t), ck.fptr(want[pos:]), 512, 0)
        rl.ref_mono_destroy(h)
        assert ck.rel_rms(got, want) <= TOL32
    assert ck.rel_rms(got, ck.direct_convolve_delayed(ir, x, 512)) <= TOL32


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_mono_latency_modes_golden(hb, mode):
    """shipped kLatencyZero / Short / Medium schemes (MonoConvolve.cpp:28-30): net delay 0 / 128 / 512."""
    ir, x, want = G["mono_ir"], G["mono_x"], G["mono_y_mode%d" % mode]
    mc = hb.MonoConvolve(len(ir), hb.LatencyMode(mode))
    mc.setResetOffset(0)
    assert mc.set(ir, len(ir), True) == 0
    tmp = np.zeros(512, np.float32)
    got = stream(lambda a, b, n: mc.process(a, tmp, b, n), x, 512)
    assert ck.rel_rms(got, want) <= TOL32


def test_mono_set_resize_and_errors(hb):
    E = hb.ConvolveError
    mc = hb.MonoConvolve(2000, False, 256)
    ir = ck.synth_ir(3000, 8)
    x = ck.synth_audio(4096, 8)
    assert mc.set(ir, 3000, False) == E.CONVOLVE_ERR_MEM_ALLOC_TOO_SMALL
    y = np.full(4096, 3.0, np.float32)
    mc.process(x, None if False else np.zeros(4096, np.float32), y, 4096)
    assert np.all(y == 3.0)            # IR longer than the allocation: process does nothing (MonoConvolve.cpp:183)
    assert mc.set(ir, 3000, True) == E.CONVOLVE_ERR_NONE
    mc.setResetOffset(0)
    mc.reset()
    mc.process(x, np.zeros(4096, np.float32), y, 4096)
    assert ck.rel_rms(y, ck.direct_convolve_delayed(ir, x, 128)) <= TOL32
    acc = np.ones(4096, np.float32)
    mc.reset()
    mc.process(x, np.zeros(4096, np.float32), acc, 4096, True)            # accumulate (MonoConvolve.cpp:167-177)
    assert ck.rel_rms(acc - 1.0, y) <= 1e-5
    with pytest.raises(RuntimeError):
        hb.MonoConvolve(1000, False, 1024, 256)
    with pytest.raises(RuntimeError):
        hb.MonoConvolve(1000, False, 16)


# ---- NToMonoConvolve / Convolver ---------------------------------------------------------------------

def test_matrix_golden_uniform(hb):
    """config-3 shape scaled down (4 -> 2, FFT 512) against the reference's time-domain sum of MonoConvolves."""
    irs, xs, want = G["matrix_irs"], G["matrix_x"], G["matrix_y"]
    fft = int(G["matrix_meta"][0])
    n_out, n_in, L = irs.shape
    cv = hb.Convolver(n_in, n_out, False, fft, maxLength=L)
    cv.setResetOffset(0)
    for o in range(n_out):
        for i in range(n_in):
            assert cv.set(i, o, irs[o, i], L, False) == 0
    got = np.zeros_like(want)
    n = xs.shape[1]
    for pos in range(0, n, 256):
        yb = np.zeros((n_out, 256), np.float32)
        cv.process(np.ascontiguousarray(xs[:, pos:pos + 256]), yb, n_in, n_out, 256)
        got[:, pos:pos + 256] = yb
    for o in range(n_out):
        assert ck.rel_rms(got[o], want[o]) <= TOL32


def test_convolver_golden_latency_short(hb):
    """Convolver(3, 2, kLatencyShort) with 17000-tap IRs through set(..., resize=true)."""
    irs, xs, want = G["conv_irs"], G["conv_x"], G["conv_y"]
    n_out, n_in, L = irs.shape
    cv = hb.Convolver(n_in, n_out, hb.kLatencyShort)
    cv.setResetOffset(0)
    assert cv.set(0, 0, irs[0, 0], L, False) == hb.CONVOLVE_ERR_MEM_ALLOC_TOO_SMALL      # default room: 16384 taps
    for o in range(n_out):
        for i in range(n_in):
            assert cv.set(i, o, irs[o, i], L, True) == 0
    got = np.zeros_like(want)
    for pos in range(0, xs.shape[1], 256):
        yb = np.zeros((n_out, 256), np.float32)
        cv.process(np.ascontiguousarray(xs[:, pos:pos + 256]), yb, n_in, n_out, 256)
        got[:, pos:pos + 256] = yb
    for o in range(n_out):
        assert ck.rel_rms(got[o], want[o]) <= TOL32


@pytest.mark.parametrize("mode", ["kLatencyShort", "kLatencyZero"])
def test_large_matrix_latency_modes_use_batches_and_match_truth(hb, mode):
    """Convolver(64, 64, LatencyMode): with 4096 pairs the small FFT parts stream enough spectra for the multi-hop batches
    (several parts accumulating into the same block, forward / inverse FFTs of a batch in one launch each, the
    register-blocked direct-form head for kLatencyZero).  Three outputs against float64 direct convolution, blocks of 1024
    samples plus a ragged call."""
    n = 64
    L = 2600
    rng = np.random.default_rng(7)
    irs = (rng.standard_normal((n, n, L)) * np.exp(-6.9 * np.arange(L) / L) * 0.1).astype(np.float32)
    xs = np.stack([ck.synth_audio(1024 * 7 + 333, 900 + i) for i in range(n)])
    cv = hb.Convolver(n, n, getattr(hb, mode))
    cv.setResetOffset(0)
    for o in range(n):
        for i in range(n):
            assert cv.set(i, o, irs[o, i], L, False) == 0
    got = np.zeros((n, xs.shape[1]), np.float32)
    pos = 0
    for m in [1024, 1024, 333, 1024, 2048, 1024, 1024]:
        yb = np.zeros((n, m), np.float32)
        cv.process(np.ascontiguousarray(xs[:, pos:pos + m]), yb, n, n, m)
        got[:, pos:pos + m] = yb
        pos += m
    assert pos == xs.shape[1]
    delay = 0 if mode == "kLatencyZero" else 128
    for o in (0, 17, 63):
        truth = sum(ck.direct_convolve_delayed(irs[o, i], xs[i], delay) for i in range(n))
        assert ck.rel_rms(got[o], truth) <= TOL32


def test_n2m_config3_reduced(hb):
    """BASELINE config 3 geometry (8 -> 1, FFT 4096) with 16384-tap IRs, against direct convolution."""
    n_in, B, L = 8, 2048, 16384
    irs = [ck.synth_ir(L, 100 + i) for i in range(n_in)]
    xs = [ck.synth_audio(B * 12, 100 + i) for i in range(n_in)]
    nm = hb.NToMonoConvolve(n_in, L, False, 2 * B)
    nm.setResetOffset(0)
    for i in range(n_in):
        assert nm.set(i, irs[i], L, False) == 0
    assert nm.set(n_in, irs[0], L, False) == hb.CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE
    out = np.zeros(B * 12, np.float32)
    tmp = np.zeros(B, np.float32)
    for pos in range(0, B * 12, B):
        nm.process([x[pos:pos + B] for x in xs], out[pos:pos + B], tmp, B, n_in)
    truth = sum(ck.direct_convolve_delayed(irs[i], xs[i], B) for i in range(n_in))
    assert ck.rel_rms(out, truth) <= TOL32
    # activeInChans < N: the remaining inputs do not contribute (NToMonoConvolve.cpp:41)
    nm.reset()
    out2 = np.zeros(B * 12, np.float32)
    for pos in range(0, B * 12, B):
        nm.process([x[pos:pos + B] for x in xs], out2[pos:pos + B], tmp, B, 3)
    truth3 = sum(ck.direct_convolve_delayed(irs[i], xs[i], B) for i in range(3))
    assert ck.rel_rms(out2, truth3) <= TOL32


def test_convolver_parallel_mode_and_double_io(hb):
    """Convolver(numIO, ...) = independent channels (Convolver.cpp:24-41); double I/O casts through float."""
    E = hb.ConvolveError
    K, B, L = 5, 128, 1000
    cv = hb.Convolver(K, False, 2 * B)
    cv.setResetOffset(0)
    irs = [ck.synth_ir(L, 200 + k) for k in range(K)]
    for k in range(K):
        assert cv.set(k, k, irs[k].astype(np.float64), L, False) == 0          # double IR overload
    assert cv.set(1, 0, irs[0], L, False) == E.CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE
    assert cv.set(K, K, irs[0], L, False) == E.CONVOLVE_ERR_OUT_CHAN_OUT_OF_RANGE
    xs = np.stack([ck.synth_audio(B * 20, 200 + k) for k in range(K)]).astype(np.float64)
    ys = np.zeros_like(xs)
    cv.process(xs, ys, K, K, xs.shape[1])
    for k in range(K):
        assert ck.rel_rms(ys[k], ck.direct_convolve_delayed(irs[k], xs[k], B)) <= TOL32
    cv.clear(2, 2, False)
    cv.reset()
    ys2 = np.zeros_like(xs)
    cv.process(xs, ys2, K, K, xs.shape[1])
    assert np.all(ys2[2] == 0) and ck.rel_rms(ys2[0], ys[0]) <= 1e-6


def test_matrix_wide_outputs_variants_agree(hb):
    """64 outputs (the config-4 tile shape: 64 rows x 64 bins per unit) on a small IR; TMA ring and
    direct-load variants must agree with each other and with direct convolution."""
    n_in, n_out, B, L = 3, 64, 256, 2000
    rng = np.random.default_rng(9)
    irs = (rng.standard_normal((n_out, n_in, L)) * np.exp(-6.9 * np.arange(L) / L)).astype(np.float32)
    xs = np.stack([ck.synth_audio(B * 16, 300 + i) for i in range(n_in)])
    res = []
    for variant in (1, 0):
        cv = hb.Convolver(n_in, n_out, False, 2 * B, maxLength=L)
        cv.matrix.tail.set_tuning(0, variant)
        cv.setResetOffset(0)
        for o in range(n_out):
            for i in range(n_in):
                assert cv.set(i, o, irs[o, i], L, False) == 0
        ys = np.zeros((n_out, xs.shape[1]), np.float32)
        cv.process(xs, ys, n_in, n_out, xs.shape[1])
        res.append(ys)
    for o in (0, 17, 63):
        truth = sum(ck.direct_convolve_delayed(irs[o, i], xs[i], B) for i in range(n_in))
        assert ck.rel_rms(res[0][o], truth) <= TOL32
    assert ck.rel_rms(res[0], res[1]) <= 1e-6


def test_linearity_and_impulse_at_config4_block_size(hb):
    """size-independent properties at config 4's geometry (FFT 8192, 64 partitions, 8 x 8 slice):
    an impulse in one input reproduces that column of IRs delayed by B, and the map is linear."""
    n_in, n_out, B, P = 8, 8, 4096, 64
    L = B * P
    rng = np.random.default_rng(10)
    cv = hb.Convolver(n_in, n_out, False, 2 * B, maxLength=L)
    cv.setResetOffset(0)
    irs = {}
    for o in range(n_out):
        for i in range(n_in):
            ir = (rng.standard_normal(L) * np.exp(-6.9 * np.arange(L) / L)).astype(np.float32)
            irs[o, i] = ir
            assert cv.set(i, o, ir, L, False) == 0
    n = L + 2 * B
    xs = np.zeros((n_in, n), np.float32)
    xs[5, 0] = 1.0
    ys = np.zeros((n_out, n), np.float32)
    cv.process(xs, ys, n_in, n_out, n)
    for o in range(n_out):
        assert np.allclose(ys[o, :B], 0, atol=1e-6)
        assert ck.rel_rms(ys[o, B:B + L], irs[o, 5]) <= TOL32
    cv.reset()
    xa = np.stack([ck.synth_audio(4 * B, 400 + i) for i in range(n_in)])
    xb = np.stack([ck.synth_audio(4 * B, 500 + i) for i in range(n_in)])
    outs = []
    for sig in (xa, xb, (xa + 0.5 * xb).astype(np.float32)):
        cv.reset()
        y = np.zeros((n_out, 4 * B), np.float32)
        cv.process(sig, y, n_in, n_out, 4 * B)
        outs.append(y.astype(np.float64))
    assert ck.rel_rms(outs[2], outs[0] + 0.5 * outs[1]) <= TOL32


def test_convolver_zero_latency_head(hb):
    """Convolver(2, 2, kLatencyZero): the direct-form head (TimeDomainConvolve.cpp:69-163) plus the four
    FFT parts give the undelayed convolution for any call sizes, including calls shorter than the head."""
    n_in, n_out, L = 2, 2, 9000
    irs = [[ck.synth_ir(L, 700 + 10 * o + i) for i in range(n_in)] for o in range(n_out)]
    xs = np.stack([ck.synth_audio(20000, 700 + i) for i in range(n_in)])
    cv = hb.Convolver(n_in, n_out, hb.kLatencyZero)
    cv.setResetOffset(0)
    for o in range(n_out):
        for i in range(n_in):
            assert cv.set(i, o, irs[o][i], L, False) == 0
    got = np.zeros((n_out, xs.shape[1]), np.float32)
    pos, k = 0, 0
    sizes = [64, 1, 100, 128, 300, 4096, 17, 1024, 2500]           # calls of 1024 samples and more take the register-blocked head
    while pos < xs.shape[1]:
        n = min(sizes[k % len(sizes)], xs.shape[1] - pos)
        yb = np.zeros((n_out, n), np.float32)
        cv.process(np.ascontiguousarray(xs[:, pos:pos + n]), yb, n_in, n_out, n)
        got[:, pos:pos + n] = yb
        pos += n
        k += 1
    for o in range(n_out):
        truth = sum(ck.direct_convolve_delayed(irs[o][i], xs[i], 0) for i in range(n_in))
        assert ck.rel_rms(got[o], truth) <= TOL32
    # an IR shorter than the head lives in the head alone
    mc = hb.MonoConvolve(100, hb.kLatencyZero)
    short = ck.synth_ir(100, 720)
    assert mc.set(short, 100, False) == 0
    y = np.zeros(5000, np.float32)
    for pos in range(0, 5000, 50):
        mc.process(xs[0][pos:pos + 50], np.zeros(50, np.float32), y[pos:pos + 50], 50)
    assert ck.rel_rms(y, ck.direct_convolve_delayed(short, xs[0][:5000], 0)) <= TOL32


def test_mono_zero_latency_against_oracle(hb):
    """custom zero-latency scheme MonoConvolve(L, true, 64, 512) against the plain-C oracle's MonoConvolve."""
    lib = ck.oracle()
    L = 3000
    ir = ck.synth_ir(L, 730)
    x = ck.synth_audio(6000, 730)
    h = lib.orc_mono_create_f32(L, 1, 64, 512, 0, 0)
    assert lib.orc_mono_set_f32(h, ck.fptr(ir), L, 0) == 0
    want = np.zeros_like(x)
    for pos in range(0, len(x), 96):
        n = min(96, len(x) - pos)
        lib.orc_mono_process_f32(h, ck.fptr(x[pos:]), ck.fptr(want[pos:]), n, 0)
    lib.orc_mono_destroy_f32(h)
    mc = hb.MonoConvolve(L, True, 64, 512)
    mc.setResetOffset(0)
    assert mc.set(ir, L, False) == 0
    got = np.zeros_like(x)
    for pos in range(0, len(x), 96):
        n = min(96, len(x) - pos)
        mc.process(x[pos:pos + n], np.zeros(n, np.float32), got[pos:pos + n], n)
    assert ck.rel_rms(got, want) <= TOL32


# ---- contracts ---------------------------------------------------------------------------------------

def test_fft_size_change_drops_every_pair(hb):
    """setFFTSize with a new size drops every loaded IR (PartitionedConvolve.cpp:145-149).  On a multi-pair engine the
    spectra left over from the old size's tiling must read as silence for the pairs that are not set again, and for the
    partitions beyond the length of a pair that is set shorter than the longest one."""
    from hisstools_library_b200.convolve import _Engine
    n_in, n_out, L = 2, 2, 4096
    e = _Engine(np.float32, 1, n_in, n_out, 1024, L, 0, 0, 0)
    e.set_reset_offset(0)
    for o in range(n_out):
        for i in range(n_in):
            e.set_ir(0, i, o, ck.synth_ir(L, 40 + 2 * o + i), L)
    xs = np.stack([ck.synth_audio(128 * 40, 40 + i) for i in range(n_in)])

    def run():
        ys = [np.zeros(xs.shape[1], np.float32) for _ in range(n_out)]
        e.process([xs[i] for i in range(n_in)], ys, xs.shape[1])
        return ys

    run()
    assert e.set_fft_size(256) == 0
    long_ir, short_ir = ck.synth_ir(128 * 5, 50), ck.synth_ir(100, 51)
    e.set_ir(0, 0, 0, long_ir, len(long_ir))                 # pair (in 0, out 0): 5 partitions
    e.set_ir(0, 1, 1, short_ir, len(short_ir))               # pair (in 1, out 1): 1 partition; the other two pairs stay dropped
    ys = run()
    assert ck.rel_rms(ys[0], ck.direct_convolve_delayed(long_ir, xs[0], 128)) <= TOL32
    assert ck.rel_rms(ys[1], ck.direct_convolve_delayed(short_ir, xs[1], 128)) <= TOL32
    e.close()


def test_process_skips_the_block_while_set_holds_the_object(hb):
    """The audio thread never waits for set / resize: while another thread replaces an impulse response, process returns
    HB_ERR_BUSY and leaves the outputs untouched -- the try-lock of MonoConvolve.cpp:181-183 / MemorySwap.h:182-185."""
    import ctypes as C
    import threading
    from hisstools_library_b200 import _abi
    lib = _abi.lib()
    mc = hb.MonoConvolve(1 << 16, hb.kLatencyShort)
    mc.setResetOffset(0)
    ir = ck.synth_ir(1 << 21, 60)
    assert mc.set(ir[:4096], 4096, False) == 0
    h = mc.matrix._h
    n = 64
    x = ck.synth_audio(n, 60)
    stop = threading.Event()
    seen = {"busy": 0, "ok": 0, "bad": 0}

    def audio_thread():
        ip = (C.c_void_p * 1)(x.ctypes.data)
        while not stop.is_set():
            y = np.full(n, 9.0, np.float32)
            op = (C.c_void_p * 1)(y.ctypes.data)
            rc = lib.hb_matrix_process(h, ip, op, n, 0)
            if rc == _abi.HB_ERR_BUSY:
                seen["busy"] += 1
                if not np.all(y == 9.0):
                    seen["bad"] += 1
            elif rc == _abi.HB_OK:
                seen["ok"] += 1
            else:
                seen["bad"] += 1

    t = threading.Thread(target=audio_thread)
    t.start()
    try:
        for k in range(6):
            # a 2M-tap IR with a resize request: allocation, upload and 100+ transforms under the object's lock
            assert mc.set(ir, len(ir) - k, True) == 0
    finally:
        stop.set()
        t.join()
    assert seen["bad"] == 0 and seen["ok"] > 0
    assert seen["busy"] > 0


@pytest.mark.parametrize("schedule", ["auto", "overlapped", "serial"])
def test_set_and_reset_of_one_pair_leave_the_other_pairs_running(hb, schedule):
    """Convolver::set / reset(in, out) on a running matrix restart ONE MonoConvolve in the reference (Convolver.cpp:88-134,
    MonoConvolve.cpp:118-152): that pair forgets its input history and, after set, answers with the new response; the other
    pairs keep playing.  Here the delay line of an input is shared by all outputs, so the pair's partitions are hidden and come
    back one per hop (hb_conv_set_ir_live).  Against the compiled reference's matrix of MonoConvolves with the same set / reset
    at the same samples: everything must agree except the one block that was already finished when the call came (still
    delivered with the old pair in it); the changed input is silent for one hop before the call, because the first frame after
    the restart shares its overlap half with the samples before it."""
    if ck.ref() is None:
        pytest.skip("compiled reference not shipped")
    lib = ck.ref()
    n_in, n_out, B, L = 3, 2, 256, 3000
    hops, t_set, t_reset = 30, 9, 19
    irs = np.stack([np.stack([ck.synth_ir(L, 2300 + 10 * o + i) for i in range(n_in)]) for o in range(n_out)])
    new_ir = ck.synth_ir(2000, 2399)
    xs = np.stack([ck.synth_audio(B * hops, 2300 + i) for i in range(n_in)])
    xs[1, (t_set - 1) * B:t_set * B] = 0                       # input 1 is silent for the hop before its pair (1, 0) is replaced
    xs[2, (t_reset - 1) * B:t_reset * B] = 0                   # input 2 before pair (2, 1) is reset
    # reference: rows of MonoConvolve(L, false, 512), one call per hop
    m = lib.ref_matrix_create(n_in, n_out, L, 2 * B, 0)
    for o in range(n_out):
        for i in range(n_in):
            lib.ref_matrix_set(m, i, o, ck.fptr(irs[o, i]), L)
    want = np.zeros((n_out, B * hops), np.float32)
    P32 = ck.c_f32p
    for k in range(hops):
        if k == t_set:
            assert lib.ref_matrix_set(m, 1, 0, ck.fptr(new_ir), len(new_ir)) == 0
        if k == t_reset:
            lib.ref_matrix_reset_pair(m, 2, 1)
        xp = (P32 * n_in)(*[xs[i, k * B:].ctypes.data_as(P32) for i in range(n_in)])
        yp = (P32 * n_out)(*[want[o, k * B:].ctypes.data_as(P32) for o in range(n_out)])
        lib.ref_matrix_process(m, xp, yp, B)
    lib.ref_matrix_destroy(m)
    # ours: the same object through the Convolver mirror, calls of one and of several hops (hops go one by one while a pair returns)
    cv = hb.Convolver(n_in, n_out, False, 2 * B, maxLength=L)
    eng = cv.matrix.tail
    eng.set_schedule(None if schedule == "auto" else schedule == "overlapped")
    cv.setResetOffset(0)
    for o in range(n_out):
        for i in range(n_in):
            assert cv.set(i, o, irs[o, i], L, False) == 0
    got = np.zeros((n_out, B * hops), np.float32)
    k = 0
    for nb in [1, 1, 3, 4, 1, 2, 4, 3, 1, 4, 6]:
        if k == t_set:
            assert cv.set(1, 0, new_ir, len(new_ir), False) == 0
        if k == t_reset:
            assert cv.reset(2, 1) == 0
        yb = np.zeros((n_out, nb * B), np.float32)
        cv.process(np.ascontiguousarray(xs[:, k * B:(k + nb) * B]), yb, n_in, n_out, nb * B)
        got[:, k * B:(k + nb) * B] = yb
        k += nb
    assert k == hops
    keep = np.ones(B * hops, bool)
    keep[t_set * B:(t_set + 1) * B] = False
    assert ck.rel_rms(got[0][keep], want[0][keep]) <= TOL32
    assert ck.rel_rms(got[1][:t_reset * B], want[1][:t_reset * B]) <= TOL32            # output 1 is not touched by the set of pair (1, 0)
    keep[:] = True
    keep[t_reset * B:(t_reset + 1) * B] = False
    assert ck.rel_rms(got[1][keep], want[1][keep]) <= TOL32
    # the pair that was replaced really answers with the new response: long after the call its old response would still ring
    assert ck.rel_rms(got[0][(t_set + 1) * B:], want[0][(t_set + 1) * B:]) <= TOL32


def test_latency_zero_convolver_set_of_one_pair_against_the_reference_object(hb):
    """The same on the shipped scheme: Convolver(2, 2, kLatencyZero) (direct-form head + four FFT sizes) against the reference's own
    Convolver object, one pair replaced at a sample where every part is at a hop boundary."""
    if ck.ref() is None:
        pytest.skip("compiled reference not shipped")
    lib = ck.ref()
    n, L, T, total, blk = 2, 20000, 16384, 49152, 512
    irs = [[ck.synth_ir(L, 2400 + 10 * o + i) for i in range(n)] for o in range(n)]
    new_ir = ck.synth_ir(12000, 2499)
    xs = np.stack([ck.synth_audio(total, 2400 + i) for i in range(n)])
    xs[0, T - 8192:T] = 0                                       # input 0 silent for the longest hop before pair (0, 1) is replaced
    h = lib.ref_conv_create(n, n, 0)
    for o in range(n):
        for i in range(n):
            assert lib.ref_conv_set_f32(h, i, o, ck.fptr(irs[o][i]), L, 1) == 0
    want = np.zeros((n, total), np.float32)
    P32 = ck.c_f32p
    for pos in range(0, total, blk):
        if pos == T:
            assert lib.ref_conv_set_f32(h, 0, 1, ck.fptr(new_ir), len(new_ir), 1) == 0
        xp = (P32 * n)(*[xs[i, pos:].ctypes.data_as(P32) for i in range(n)])
        yp = (P32 * n)(*[want[o, pos:].ctypes.data_as(P32) for o in range(n)])
        lib.ref_conv_process_f32(h, xp, yp, n, n, blk)
    lib.ref_conv_destroy(h)
    cv = hb.Convolver(n, n, hb.kLatencyZero)
    cv.setResetOffset(0)
    for o in range(n):
        for i in range(n):
            assert cv.set(i, o, irs[o][i], L, True) == 0
    got = np.zeros((n, total), np.float32)
    for pos in range(0, total, blk):
        if pos == T:
            assert cv.set(0, 1, new_ir, len(new_ir), True) == 0
        yb = np.zeros((n, blk), np.float32)
        cv.process(np.ascontiguousarray(xs[:, pos:pos + blk]), yb, n, n, blk)
        got[:, pos:pos + blk] = yb
    assert ck.rel_rms(got[0], want[0]) <= TOL32                                        # output 0: untouched by the set of pair (0, 1)
    assert ck.rel_rms(got[1][:T], want[1][:T]) <= TOL32
    assert ck.rel_rms(got[1][T + 8192:], want[1][T + 8192:]) <= TOL32                  # after the blocks that were already finished


@pytest.mark.parametrize("ins,outs,groups,B,L", [(1, 1, 1, 512, 4096), (8, 1, 1, 2048, 40000), (3, 2, 2, 128, 1000), (1, 1, 1, 1024, 65536)])
def test_hop_batches_on_small_engines(hb, ins, outs, groups, B, L):
    """Launch-latency-bound engines (BASELINE configs 1-3 and their like) take the hops of a multi-block call as one set of
    launches -- forward FFTs of all hops, multiply-accumulate over hop x bin, inverse FFTs of all hops -- instead of one
    launch sequence per hop: calls of 64, 5, 2 ... hops mixed with single blocks and ragged calls against the same engine
    fed hop by hop (summation order only) and, for the first output, the reference / float64 direct convolution."""
    from hisstools_library_b200.convolve import _Engine
    irs = [[[ck.synth_ir(L, 2100 + 100 * g + 10 * o + i) for i in range(ins)] for o in range(outs)] for g in range(groups)]
    calls = [64 * B, B, 5 * B, 100, B - 100, 2 * B, 70 * B, B, 3 * B + 17, B - 17, 9 * B]
    n = sum(calls)
    xs = np.stack([ck.synth_audio(n, 2100 + r) for r in range(groups * ins)])
    res = {}
    for batches in (True, False):
        e = _Engine(np.float32, groups, ins, outs, 2 * B, L, 0, 0, 0)
        e.set_multi_hop(batches)
        e.set_reset_offset(0)
        for g in range(groups):
            for o in range(outs):
                for i in range(ins):
                    e.set_ir(g, i, o, irs[g][o][i], L)
        y = np.zeros((groups * outs, n), np.float32)
        pos = 0
        for m in calls:
            yo = [np.zeros(m, np.float32) for _ in range(groups * outs)]
            e.process([np.ascontiguousarray(xs[r, pos:pos + m]) for r in range(groups * ins)], yo, m)
            for r in range(groups * outs):
                y[r, pos:pos + m] = yo[r]
            pos += m
        res[batches] = y
        e.close()
    for r in range(groups * outs):
        assert ck.rel_rms(res[True][r], res[False][r]) <= 2e-6
    truth = sum(ck.direct_convolve_delayed_fft(irs[0][0][i], xs[i], B) for i in range(ins))
    assert ck.rel_rms(res[True][0], truth) <= TOL32


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("ins,outs,groups,B,L", [(1, 1, 1, 512, 4096), (1, 1, 1, 1024, 65536), (8, 1, 1, 2048, 131072), (3, 1, 2, 256, 5000),
                                                 (1, 1, 1, 64, 64), (1, 1, 1, 128, 256), (2, 3, 1, 512, 3000), (20, 1, 1, 128, 1024)])
def test_fused_hops_overlap_between_calls(hb, dtype, ins, outs, groups, B, L):
    """Back-to-back single-block device calls on a fused engine (hb_conv_set_hop_overlap): mode 0 keeps every hop behind the
    previous one; mode 2 (the caller's rows are complete when a call is made) and the engine's own stream in mode 1 let hop t+1
    transform its frame while hop t still multiplies, ordered by the counters of hb_conv_fused.cuh.  300 calls with distinct
    input and output rows per call: the samples are bit-identical in all three, and what the reference / float64 direct
    convolution gives; BASELINE configs 1-3, one and two partitions, several outputs, more inputs than ranks in the cluster."""
    import torch
    from hisstools_library_b200.convolve import _Engine
    if dtype == np.float64 and ins * L > 500000:
        pytest.skip("not a fused engine in double")
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    tol = TOL32 if dtype == np.float32 else TOL64
    hops = 300
    n = hops * B
    irs = [[[ck.synth_ir(L, 3100 + 100 * g + 10 * o + i).astype(dtype) for i in range(ins)] for o in range(outs)] for g in range(groups)]
    xs = np.stack([ck.synth_audio(n, 3100 + r) for r in range(groups * ins)]).astype(dtype)
    x = torch.from_numpy(xs).cuda()
    res = {}
    for mode, own in ((0, False), (2, False), (1, True), (1, False)):
        e = _Engine(dtype, groups, ins, outs, 2 * B, L, 0, 0, 0)
        e.set_reset_offset(0)
        e.set_hop_overlap(mode)
        for g in range(groups):
            for o in range(outs):
                for i in range(ins):
                    e.set_ir(g, i, o, irs[g][o][i], L)
        y = torch.zeros((groups * outs, n), dtype=tdt, device="cuda")
        torch.cuda.synchronize()
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            for h in range(hops):
                e.process_device(x.data_ptr() + h * B * x.element_size(), n, y.data_ptr() + h * B * y.element_size(), n, B, False,
                                 0 if own else stream.cuda_stream)
            if h == hops - 1:
                e.join(stream.cuda_stream)
        torch.cuda.synchronize()
        assert e.schedule == "fused"
        res[(mode, own)] = y.cpu().numpy()
        e.close()
    base = res[(0, False)]
    for key in ((2, False), (1, True), (1, False)):
        assert np.array_equal(res[key], base), key
    for g in range(groups):
        for o in range(outs):
            truth = sum(ck.direct_convolve_delayed_fft(irs[g][o][i], xs[g * ins + i], B) for i in range(ins))
            assert ck.rel_rms(base[g * outs + o], truth) <= tol * (1 if dtype == np.float32 else 10)


def test_fused_hops_overlap_ragged_calls_and_resets(hb):
    """Mode 2 with everything that interrupts a chain of overlapping hops: ragged calls through the staging rows, calls of several
    blocks (hop batches), a reset, a new impulse response on the running engine, host-pointer calls in between -- bit-identical to
    the same sequence in mode 0."""
    import torch
    from hisstools_library_b200.convolve import _Engine
    B, L, ins = 256, 4000, 2
    irs = [ck.synth_ir(L, 3300 + i) for i in range(ins)]
    irs2 = [ck.synth_ir(L // 2, 3400 + i) for i in range(ins)]
    calls = [B] * 20 + [100, B - 100] + [B] * 5 + [3 * B] + [B] * 7 + [17] + [B] * 9 + [2 * B - 17] + [B] * 30
    n = sum(calls)
    xs = np.stack([ck.synth_audio(n, 3300 + r) for r in range(ins)])
    x = torch.from_numpy(xs).cuda()
    res = {}
    for mode in (0, 2):
        e = _Engine(np.float32, 1, ins, 1, 2 * B, L, 0, 0, 0)
        e.set_reset_offset(0)
        e.set_hop_overlap(mode)
        for i in range(ins):
            e.set_ir(0, i, 0, irs[i], L)
        y = torch.zeros((1, n), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        stream = torch.cuda.Stream()
        pos = 0
        with torch.cuda.stream(stream):
            for k, m in enumerate(calls):
                if k == 30:
                    torch.cuda.synchronize()
                    e.reset()
                if k == 45:
                    torch.cuda.synchronize()
                    e.set_ir_live(0, 1, 0, irs2[1], L // 2)
                if k in (50, 51):
                    torch.cuda.synchronize()
                    yo = [np.zeros(m, np.float32)]
                    e.process([np.ascontiguousarray(xs[r, pos:pos + m]) for r in range(ins)], yo, m)
                    y[0, pos:pos + m] = torch.from_numpy(yo[0]).cuda()
                    torch.cuda.synchronize()
                else:
                    e.process_device(x.data_ptr() + pos * 4, n, y.data_ptr() + pos * 4, n, m, False, stream.cuda_stream)
                pos += m
        torch.cuda.synchronize()
        res[mode] = y.cpu().numpy()
        e.close()
    assert np.array_equal(res[0], res[2])


@pytest.mark.parametrize("scheme,ins,outs,L", [((False, 1024, 0, 0, 0), 1, 1, 20000), ((False, 512, 0, 0, 0), 4, 2, 6000),
                                               ((False, 256, 1024, 4096, 16384), 2, 1, 40000), ((True, 256, 1024, 4096, 16384), 1, 1, 30000)])
def test_matrix_device_calls_overlap_between_calls(hb, scheme, ins, outs, L):
    """hb_matrix_set_hop_overlap: device calls of one block on a matrix -- a uniform MonoConvolve / Convolver, the reference's
    short-latency scheme (three fixed parts and the tail, every one a fused engine here) and the zero-latency one with its
    direct-form head between the parts -- with the matrix's own stream (mode 1), any stream (mode 2) and no overlap (mode 0):
    bit-identical samples, equal to float64 direct convolution."""
    import torch
    from hisstools_library_b200.convolve import _Matrix
    blk = 256
    hops = 200
    n = hops * blk
    irs = [[ck.synth_ir(L, 3500 + 10 * o + i) for i in range(ins)] for o in range(outs)]
    xs = np.stack([ck.synth_audio(n, 3500 + r) for r in range(ins)])
    x = torch.from_numpy(xs).cuda()
    res = {}
    for mode, own in ((0, False), (2, False), (1, True)):
        m = _Matrix(1, ins, outs, L, scheme, np.float32, 0)
        m.setResetOffset(0)
        m.set_hop_overlap(mode)
        for o in range(outs):
            for i in range(ins):
                assert int(m.set(0, i, o, irs[o][i], L, True)) == 0
        y = torch.zeros((outs, n), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            for h in range(hops):
                assert m.process_device(x.data_ptr() + h * blk * 4, n, y.data_ptr() + h * blk * 4, n, blk, False, 0 if own else stream.cuda_stream)
        torch.cuda.synchronize()
        res[(mode, own)] = y.cpu().numpy()
        m.close()
    base = res[(0, False)]
    assert np.array_equal(res[(2, False)], base)
    assert np.array_equal(res[(1, True)], base)
    lat = 0 if scheme[0] else scheme[1] // 2
    for o in range(outs):
        truth = sum(np.convolve(xs[i].astype(np.float64), irs[o][i].astype(np.float64))[:n] for i in range(ins))
        truth = np.concatenate([np.zeros(lat), truth])[:n]
        assert ck.rel_rms(base[o], truth) <= TOL32


def test_fused_hops_fed_with_their_own_output_keep_stream_order(hb):
    """An engine in a feedback loop -- the output rows of call t are the input rows of call t+1, nothing else in between -- relies on
    stream order for its input: such calls are recognised by their row addresses and keep the strict order even where overlapping
    hops are allowed (mode 2, and the engine's own stream in mode 1).  Bit-identical to mode 0."""
    import torch
    from hisstools_library_b200.convolve import _Engine
    B, L, hops = 256, 1500, 60
    ir = (ck.synth_ir(L, 3600) * 0.2).astype(np.float32)
    first = ck.synth_audio(B, 3600)
    res = {}
    for mode, own in ((0, False), (2, False), (1, True)):
        e = _Engine(np.float32, 1, 1, 1, 2 * B, L, 0, 0, 0)
        e.set_reset_offset(0)
        e.set_hop_overlap(mode)
        e.set_ir(0, 0, 0, ir, L)
        ring = torch.zeros((3, B), dtype=torch.float32, device="cuda")
        ring[0] = torch.from_numpy(first).cuda()
        keep = torch.zeros((hops, B), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            for t in range(hops):
                e.process_device(ring[t % 3].data_ptr(), B, ring[(t + 1) % 3].data_ptr(), B, B, False, 0 if own else stream.cuda_stream)
                if t % 7 == 6:
                    # now and then look at the rows (on the engine's own stream the caller has to wait for it first)
                    torch.cuda.synchronize()
                    keep[t] = ring[(t + 1) % 3]
                    torch.cuda.synchronize()
        torch.cuda.synchronize()
        res[(mode, own)] = (keep.cpu().numpy(), ring.cpu().numpy())
        e.close()
    for key in ((2, False), (1, True)):
        assert np.array_equal(res[key][0], res[(0, False)][0])
        assert np.array_equal(res[key][1], res[(0, False)][1])
    assert np.abs(res[(0, False)][1]).max() > 0


@pytest.mark.parametrize("ins,B,L", [(1, 1024, 65536), (8, 2048, 131072), (1, 512, 4096)])
def test_fused_hops_strict_and_overlapping_hops_interleaved(hb, ins, B, L):
    """Strict hops directly behind overlapping ones and the other way round, back to back in one stream: every few calls something
    ends the chain without enqueuing anything (hb_conv_join, a mode change, a getter that takes the lock), so the next hop runs in the
    strict order while its predecessors may still be in flight (its products of partitions >= 2 then wait as well).  400 calls on
    engines with 8 to 64 partitions on clusters of up to 16: bit-identical to mode 0."""
    import torch
    from hisstools_library_b200.convolve import _Engine
    hops = 400
    n = hops * B
    irs = [ck.synth_ir(L, 3700 + i) for i in range(ins)]
    xs = np.stack([ck.synth_audio(n, 3700 + r) for r in range(ins)])
    x = torch.from_numpy(xs).cuda()
    res = {}
    for mode in (0, 2):
        e = _Engine(np.float32, 1, ins, 1, 2 * B, L, 0, 0, 0)
        e.set_reset_offset(0)
        e.set_hop_overlap(mode)
        for i in range(ins):
            e.set_ir(0, i, 0, irs[i], L)
        y = torch.zeros((1, n), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            for h in range(hops):
                if mode and h % 3 == 2:
                    e.join(stream.cuda_stream)
                if mode and h % 50 == 25:
                    e.set_hop_overlap(0)
                if mode and h % 50 == 30:
                    e.set_hop_overlap(2)
                if mode and h % 7 == 0:
                    assert e.tail_streams >= 0
                e.process_device(x.data_ptr() + h * B * 4, n, y.data_ptr() + h * B * 4, n, B, False, stream.cuda_stream)
        torch.cuda.synchronize()
        assert e.schedule == "fused"
        res[mode] = y.cpu().numpy()
        e.close()
    assert np.array_equal(res[0], res[2])
    truth = sum(ck.direct_convolve_delayed_fft(irs[i], xs[i], B) for i in range(ins))
    assert ck.rel_rms(res[0][0], truth) <= TOL32


@pytest.mark.parametrize("dtype,groups,ins,outs,B", [(np.float64, 16, 1, 1, 8192), (np.float32, 1, 64, 8, 4096), (np.float32, 1, 6, 40, 4096)])
def test_host_calls_with_large_rows(hb, dtype, groups, ins, outs, B):
    """Host-pointer calls of one block that carry a MiB each way (rows copied by the helper threads, calls pipelined behind the
    one-hop latency): config 5's and config 4's row shapes with short IRs, and a matrix whose outputs but not its inputs reach
    the size, against float64 direct convolution; accumulate, null rows and a ragged call in between."""
    from hisstools_library_b200.convolve import _Engine
    L = 3 * B + 100
    tol = TOL32 if dtype == np.float32 else TOL64
    hops = 7
    n = hops * B
    irs = [[[ck.synth_ir(L, 3800 + 100 * g + 10 * o + i).astype(dtype) for i in range(ins)] for o in range(outs)] for g in range(groups)]
    xs = np.stack([ck.synth_audio(n, 3800 + r) for r in range(groups * ins)]).astype(dtype)
    xs[1] = 0                                                # this row is passed as a null pointer
    e = _Engine(dtype, groups, ins, outs, 2 * B, L, 0, 0, 0)
    e.set_reset_offset(0)
    for g in range(groups):
        for o in range(outs):
            for i in range(ins):
                e.set_ir(g, i, o, irs[g][o][i], L)
    y = np.zeros((groups * outs, n), dtype)
    calls = [B, B, 1000, B - 1000, B, B, B, B]
    pos = 0
    for k, m in enumerate(calls):
        rows_in = [None if r == 1 else np.ascontiguousarray(xs[r, pos:pos + m]) for r in range(groups * ins)]
        yo = [np.full(m, 0.5, dtype) if k == 4 else np.zeros(m, dtype) for _ in range(groups * outs)]
        e.process(rows_in, yo, m, accumulate=(k == 4))
        for r in range(groups * outs):
            y[r, pos:pos + m] = yo[r] - (0.5 if k == 4 else 0.0)
        pos += m
    e.close()
    for g in range(groups):
        for o in range(0, outs, max(1, outs // 4)):
            truth = sum(ck.direct_convolve_delayed_fft(irs[g][o][i], xs[g * ins + i], B) for i in range(ins))
            assert ck.rel_rms(y[g * outs + o], truth) <= tol * (1 if dtype == np.float32 else 10) * (3 if dtype == np.float32 else 1)
