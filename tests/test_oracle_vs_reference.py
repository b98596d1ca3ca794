"""Pins oracle/ (our plain-C restatement) against the UNMODIFIED reference compiled in place
(oracle/_ref).  CPU only.  Skips when oracle/_ref is absent (it is built here by
`make -C oracle ref`, and shipped prebuilt to the GPU box).

Tolerances: both sides are independent float implementations of the same mathematics, each about
3e-7 (float) / 1e-16 (double) relative RMS from exact, so they must agree to 2e-6 / 1e-13.
"""
import ctypes as C

import numpy as np
import pytest

import checkers as ck

pytestmark = pytest.mark.skipif(ck.ref() is None, reason="oracle/_ref not built")

TOL = {np.float32: 2e-6, np.float64: 1e-13}


def _setup_pair(dtype, log2n):
    suf = ck.SUF[np.dtype(dtype)]
    r = getattr(ck.ref(), "ref_fft_setup" + suf)(max(log2n, 3))
    o = getattr(ck.oracle(), "orc_fft_setup_create" + suf)(max(log2n, 1))
    return suf, r, o


def _free_pair(suf, r, o):
    getattr(ck.ref(), "ref_fft_setup_free" + suf)(r)
    getattr(ck.oracle(), "orc_fft_setup_destroy" + suf)(o)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("log2n", list(range(0, 15)) + [17])
@pytest.mark.parametrize("op", ["fft", "ifft", "rfft", "rifft"])
def test_inplace_transforms(dtype, log2n, op):
    """FFT_Tester-shaped sweep ("- Test/FFT_Tester/FFT_Tester/main.cpp":87-140) with a result check."""
    if op in ("rfft", "rifft") and log2n == 0:
        pytest.skip("real transforms start at log2n 1")
    rng = np.random.default_rng(log2n * 7 + 1)
    n = 1 << log2n
    planes = n if op in ("fft", "ifft") else max(n >> 1, 1)
    re = rng.uniform(-1, 1, planes).astype(dtype)
    im = rng.uniform(-1, 1, planes).astype(dtype)
    suf, rs, os_ = _setup_pair(dtype, log2n)
    r_re, r_im, o_re, o_im = re.copy(), im.copy(), re.copy(), im.copy()
    getattr(ck.ref(), "ref_%s%s" % (op, suf))(rs, ck.fptr(r_re), ck.fptr(r_im), log2n)
    getattr(ck.oracle(), "orc_%s%s" % (op, suf))(os_, ck.fptr(o_re), ck.fptr(o_im), log2n)
    _free_pair(suf, rs, os_)
    err = ck.rel_rms(np.concatenate([o_re, o_im]), np.concatenate([r_re, r_im]))
    assert err < TOL[dtype], err


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_conventions_against_numpy(dtype):
    """2x DFT, DC in realp[0], Nyquist in imagp[0], rifft(rfft(x)) = 2N x."""
    log2n = 10
    n = 1 << log2n
    x = np.random.default_rng(3).uniform(-1, 1, n).astype(dtype)
    suf, rs, os_ = _setup_pair(dtype, log2n)
    re = np.zeros(n // 2, dtype)
    im = np.zeros(n // 2, dtype)
    getattr(ck.oracle(), "orc_rfft_real" + suf)(os_, ck.fptr(x), ck.fptr(re), ck.fptr(im), n, log2n)
    want = 2 * np.fft.rfft(x.astype(np.float64))
    got = re.astype(np.float64) + 1j * im
    assert abs(got[0].real - want[0].real) < 1e-3 and abs(got[0].imag - want[n // 2].real) < 1e-3
    tol = 1e-5 if dtype == np.float32 else 1e-12
    assert ck.rel_rms(np.concatenate([got[1:].real, got[1:].imag]),
                      np.concatenate([want[1:-1].real, want[1:-1].imag])) < tol
    back = np.zeros(n, dtype)
    getattr(ck.oracle(), "orc_rifft_real" + suf)(os_, ck.fptr(re), ck.fptr(im), ck.fptr(back), log2n)
    assert ck.rel_rms(back, 2 * n * x.astype(np.float64)) < tol
    _free_pair(suf, rs, os_)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("log2n,in_length", [(3, 8), (5, 32), (5, 17), (6, 1), (8, 100), (8, 255), (10, 2000), (4, 16)])
def test_rfft_real_with_in_length(dtype, log2n, in_length):
    """out-of-place rfft with zero padding and the odd-sample rule (HISSTools_FFT_Core.h:1258-1287)."""
    n = 1 << log2n
    x = np.random.default_rng(in_length).uniform(-1, 1, max(in_length, 1)).astype(dtype)
    suf, rs, os_ = _setup_pair(dtype, log2n)
    out = []
    for lib, name, s in ((ck.ref(), "ref_rfft_real", rs), (ck.oracle(), "orc_rfft_real", os_)):
        re = np.full(n // 2, 7, dtype)
        im = np.full(n // 2, 7, dtype)
        getattr(lib, name + suf)(s, ck.fptr(x), ck.fptr(re), ck.fptr(im), in_length, log2n)
        out.append(np.concatenate([re, im]))
    _free_pair(suf, rs, os_)
    assert ck.rel_rms(out[1], out[0]) < TOL[dtype]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("log2n", range(1, 18))
def test_zip_unzip_exact(dtype, log2n):
    """the reference's only known-answer test on this path: integer ramps through unzip/zip
    ("- Test/FFT_Tester/FFT_Tester/main.cpp":201-250) -- exact."""
    n = 1 << log2n
    ramp = np.arange(n).astype(dtype)
    suf = ck.SUF[np.dtype(dtype)]
    for lib, pre in ((ck.ref(), "ref_"), (ck.oracle(), "orc_")):
        re = np.zeros(n // 2, dtype)
        im = np.zeros(n // 2, dtype)
        getattr(lib, pre + "unzip" + suf)(ck.fptr(ramp), ck.fptr(re), ck.fptr(im), log2n)
        assert np.array_equal(re, ramp[0::2]) and np.array_equal(im, ramp[1::2])
        back = np.zeros(n, dtype)
        getattr(lib, pre + "zip" + suf)(ck.fptr(re), ck.fptr(im), ck.fptr(back), log2n)
        assert np.array_equal(back, ramp)


CASES = [
    # fft, ir_len, block, max_len, offset, length, reset_offset
    (1024, 4096, 512, None, 0, 0, 0),          # BASELINE config 1 shape (as a bare PartitionedConvolve)
    (2048, 65536, 1024, None, 0, 0, 0),        # BASELINE config 2
    (64, 100, 32, None, 0, 0, 0),              # IR not a multiple of B
    (64, 10, 32, None, 0, 0, 0),               # IR shorter than one partition
    (256, 1000, 77, None, 0, 0, 0),            # ragged call size
    (256, 1000, 1, None, 0, 0, 0),             # one sample per call
    (256, 1000, 129, None, 0, 0, 17),          # non-zero reset phase
    (512, 4000, 300, 4000, 1000, 1500, 0),     # offset / length slicing
    (512, 1281, 256, 1000, 0, 0, 0),           # maxLength rounding: 1000 -> 1024, error 4 + truncation
    (32, 64, 16, None, 0, 0, 0),               # minimum FFT size
]


@pytest.mark.parametrize("fft,ir_len,block,max_len,offset,length,reset_offset", CASES)
def test_pconv_float(fft, ir_len, block, max_len, offset, length, reset_offset):
    ir = ck.synth_ir(ir_len, 1)
    hops = (ir_len * 2) // (fft // 2) + 6
    x = ck.synth_audio(hops * (fft // 2) + 13, 2)
    yr, er = ck.ref_pconv_run(fft, ir, x, block, max_len, offset, length, reset_offset)
    yo, eo = ck.oracle_pconv_run(fft, ir, x, block, max_len, offset, length, reset_offset)
    assert er == eo
    assert ck.rel_rms(yo, yr) < 2e-6
    # and both equal the true convolution delayed by B
    eff = ir[offset:]
    if length:
        eff = eff[:length]
    cap = max_len if max_len is not None else ir_len
    half = fft // 2
    cap = -(-cap // half) * half
    eff = eff[:cap]
    truth = ck.direct_convolve_delayed(eff, x, half)
    assert ck.rel_rms(yr, truth) < 2e-6
    assert ck.rel_rms(yo, truth) < 2e-6


def test_pconv_chunking_is_bit_identical():
    """reference property (SURVEY B): output does not depend on how process() calls are chunked."""
    ir = ck.synth_ir(3000, 5)
    x = ck.synth_audio(9000, 5)
    a, _ = ck.oracle_pconv_run(256, ir, x, 128)
    rng = np.random.default_rng(0)
    lib = ck.oracle()
    h = lib.orc_pconv_create_f32(256, len(ir), 0, 0)
    lib.orc_pconv_set_reset_offset_f32(h, 0)
    lib.orc_pconv_set_f32(h, ck.fptr(ir), len(ir))
    b = np.zeros_like(x)
    pos = 0
    while pos < len(x):
        n = min(int(rng.integers(1, 700)), len(x) - pos)
        lib.orc_pconv_process_f32(h, ck.fptr(x[pos:]), ck.fptr(b[pos:]), n)
        pos += n
    lib.orc_pconv_destroy_f32(h)
    assert np.array_equal(a, b)


def test_pconv_no_ir_and_errors():
    lib, rl = ck.oracle(), ck.ref()
    x = ck.synth_audio(64)
    for create, proc, destroy, setfft, setlen in (
            (lib.orc_pconv_create_f32, lib.orc_pconv_process_f32, lib.orc_pconv_destroy_f32,
             lib.orc_pconv_set_fft_size_f32, lib.orc_pconv_set_length_f32),
            (rl.ref_pconv_create, rl.ref_pconv_process, rl.ref_pconv_destroy,
             rl.ref_pconv_set_fft_size, rl.ref_pconv_set_length)):
        h = create(1024, 4096, 0, 0)
        y = np.full(64, 5, np.float32)
        assert proc(h, ck.fptr(x), ck.fptr(y), 64) == 0
        assert np.all(y == 5)                      # out untouched (PartitionedConvolve.cpp:262-263)
        assert setfft(h, 16) == 11                 # below 2^5
        assert setfft(h, 2048) == 11               # above max
        assert setfft(h, 500) == 12                # not a power of two (rounds up to 512)
        assert setfft(h, 512) == 0
        assert setlen(h, 5000) == 7
        assert setlen(h, 4096) == 0
        destroy(h)


def test_pconv_fft_size_change_invalidates_ir():
    lib, rl = ck.oracle(), ck.ref()
    ir = ck.synth_ir(600)
    x = ck.synth_audio(2048)
    outs = []
    for pre, L, suf in (("orc_pconv_", lib, "_f32"), ("ref_pconv_", rl, "")):
        g = lambda n: getattr(L, pre + n + suf)
        h = g("create")(1024, 2048, 0, 0)
        g("set_reset_offset")(h, 0)
        g("set")(h, ck.fptr(ir), len(ir))
        y = np.zeros_like(x)
        assert g("process")(h, ck.fptr(x), ck.fptr(y), 512) == 1
        assert g("set_fft_size")(h, 256) == 0
        assert g("process")(h, ck.fptr(x), ck.fptr(y), 512) == 0   # IR gone until set() again
        g("set")(h, ck.fptr(ir), len(ir))
        y = np.zeros_like(x)
        assert g("process")(h, ck.fptr(x), ck.fptr(y), len(x)) == 1
        outs.append(y)
        g("destroy")(h)
    assert ck.rel_rms(outs[0], outs[1]) < 2e-6


def test_restated_float_matches_real_class_then_double_is_trusted():
    """the T-generic restatement over the reference FFT must reproduce the real float
    class before its double instantiation is used as the C5 oracle; then our C oracle in double
    must agree with it to 1e-13."""
    rl, lib = ck.ref(), ck.oracle()
    ir = ck.synth_ir(5000, 9)
    x = ck.synth_audio(256 * 60, 9)
    yr, _ = ck.ref_pconv_run(512, ir, x, 256)
    h = rl.ref_restated_create_f32(512)
    rl.ref_restated_set_f32(h, ck.fptr(ir), len(ir))
    yf = np.zeros_like(x)
    rl.ref_restated_process_f32(h, ck.fptr(x), ck.fptr(yf), len(x))
    rl.ref_restated_destroy_f32(h)
    assert ck.rel_rms(yf, yr) < 1e-7
    ird, xd = ir.astype(np.float64), x.astype(np.float64)
    h = rl.ref_restated_create_f64(512)
    rl.ref_restated_set_f64(h, ck.fptr(ird), len(ird))
    yd = np.zeros_like(xd)
    rl.ref_restated_process_f64(h, ck.fptr(xd), ck.fptr(yd), len(xd))
    rl.ref_restated_destroy_f64(h)
    yo, _ = ck.oracle_pconv_run(512, ird, xd, 256, dtype=np.float64)
    truth = ck.direct_convolve_delayed(ird, xd, 256)
    assert ck.rel_rms(yd, truth) < 1e-13
    assert ck.rel_rms(yo, yd) < 1e-13


MONO = [
    # zero_latency, sizes, ir_len, delay
    (0, (1024, 0, 0, 0), 4096, 512),                  # BASELINE config 1: MonoConvolve(4096,false,1024)
    (0, (256, 1024, 4096, 16384), 40000, 128),        # kLatencyShort
    (0, (1024, 4096, 16384, 0), 40000, 512),          # kLatencyMedium
    (1, (256, 1024, 4096, 16384), 40000, 0),          # kLatencyZero
    (1, (64, 256, 0, 0), 3000, None),                 # reference quirk, see the test body
    (0, (64, 128, 0, 0), 50, 32),                     # IR ends inside the first fixed part
]


@pytest.mark.parametrize("zero,sizes,ir_len,delay", MONO)
def test_mono_partition_schemes(zero, sizes, ir_len, delay):
    rl, lib = ck.ref(), ck.oracle()
    ir = ck.synth_ir(ir_len, 3)
    x = ck.synth_audio(ir_len * 2 + 5000, 3)
    rh = rl.ref_mono_create_custom(ir_len, zero, *sizes)
    rl.ref_mono_set_reset_offset(rh, 0)
    assert rl.ref_mono_set(rh, ck.fptr(ir), len(ir), 1) == 0
    rl.ref_mono_set_reset_offset(rh, 0)
    oh = lib.orc_mono_create_f32(ir_len, zero, *sizes)
    assert lib.orc_mono_set_f32(oh, ck.fptr(ir), len(ir), 1) == 0
    yr, yo, tmp = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
    pos, block = 0, 333
    while pos < len(x):
        n = min(block, len(x) - pos)
        rl.ref_mono_process(rh, ck.fptr(x[pos:]), ck.fptr(tmp), ck.fptr(yr[pos:]), n, 0)
        lib.orc_mono_process_f32(oh, ck.fptr(x[pos:]), ck.fptr(yo[pos:]), n, 0)
        pos += n
    rl.ref_mono_destroy(rh)
    lib.orc_mono_destroy_f32(oh)
    if delay is None:
        # zeroLatency with fewer than four sizes: the part after the head runs with
        # `accumulate || mPart1` / `|| mPart2` == false (MonoConvolve.cpp:196-198) and OVERWRITES the
        # head's output, so taps [0, A/2) are lost.  The drop-in reproduces the reference, not the maths.
        lost = ir.copy()
        lost[: sizes[0] // 2] = 0
        truth = ck.direct_convolve_delayed(lost, x, 0)
    else:
        truth = ck.direct_convolve_delayed(ir, x, delay)
    assert ck.rel_rms(yr, truth) < 2e-6
    assert ck.rel_rms(yo, yr) < 2e-6


def test_mono_invalid_sizes_and_alloc_errors():
    rl, lib = ck.ref(), ck.oracle()
    assert rl.ref_mono_create_custom(1000, 0, 16, 0, 0, 0) is None
    assert lib.orc_mono_create_f32(1000, 0, 16, 0, 0, 0) is None
    assert rl.ref_mono_create_custom(1000, 0, 1024, 256, 0, 0) is None
    assert lib.orc_mono_create_f32(1000, 0, 1024, 256, 0, 0) is None
    ir = ck.synth_ir(20000)
    rh = rl.ref_mono_create_custom(16384, 0, 1024, 0, 0, 0)
    oh = lib.orc_mono_create_f32(16384, 0, 1024, 0, 0, 0)
    assert rl.ref_mono_set(rh, ck.fptr(ir), len(ir), 0) == 4
    assert lib.orc_mono_set_f32(oh, ck.fptr(ir), len(ir), 0) == 4
    # over-long IR without resize: process() is silent on both (MonoConvolve.cpp:183)
    x = ck.synth_audio(2048)
    yr, yo, tmp = np.full(2048, 3, np.float32), np.full(2048, 3, np.float32), np.zeros(2048, np.float32)
    rl.ref_mono_process(rh, ck.fptr(x), ck.fptr(tmp), ck.fptr(yr), 2048, 0)
    lib.orc_mono_process_f32(oh, ck.fptr(x), ck.fptr(yo), 2048, 0)
    assert np.all(yr == 3) and np.all(yo == 3)
    rl.ref_mono_destroy(rh)
    lib.orc_mono_destroy_f32(oh)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", range(5))
@pytest.mark.parametrize("n1,n2", [(1, 1), (1, 9), (1000, 300), (300, 1000), (64, 64), (7, 2), (513, 512)])
def test_spectral_convolve(dtype, mode, n1, n2):
    rs, lib = ck.ref_spectral(), ck.oracle()
    suf = ck.SUF[np.dtype(dtype)]
    rng = np.random.default_rng(n1 * 31 + n2)
    a = rng.uniform(-1, 1, n1).astype(dtype)
    b = rng.uniform(-1, 1, n2).astype(dtype)
    outs = []
    for fn in (getattr(rs, "ref_spectral_convolve" + suf), getattr(lib, "orc_spectral_convolve" + suf)):
        y = np.zeros(n1 + n2 + 8, dtype)
        size = fn(ck.fptr(y), ck.fptr(a), n1, ck.fptr(b), n2, mode, 32768)
        outs.append((size, y))
    assert outs[0][0] == outs[1][0] > 0
    assert ck.rel_rms(outs[1][1], outs[0][1]) < (2e-6 if dtype == np.float32 else 1e-13)
    if mode == 0:
        truth = np.convolve(a.astype(np.float64), b.astype(np.float64))
        assert ck.rel_rms(outs[1][1][: n1 + n2 - 1], truth) < (2e-6 if dtype == np.float32 else 1e-13)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n1,n2", [(1, 1), (1, 9), (9, 1), (1000, 300), (300, 1000), (64, 64), (7, 2), (2, 7), (513, 512), (2000, 1999)])
def test_spectral_correlate_and_complex(dtype, n1, n2):
    """correlate (real, all edge modes) and the complex-input operations (all edge modes where the reference's
    reads stay inside its transform; Linear / Fold / FoldRepeat always do) -- oracle against the compiled reference,
    and Linear against direct float64 correlation / convolution."""
    rs, lib = ck.ref_spectral(), ck.oracle()
    if rs is None:
        pytest.skip("compiled reference not available")
    suf = ck.SUF[np.dtype(dtype)]
    tol = 2e-6 if dtype == np.float32 else 1e-13
    rng = np.random.default_rng(n1 * 37 + n2)
    a = rng.uniform(-1, 1, n1).astype(dtype)
    b = rng.uniform(-1, 1, n2).astype(dtype)
    for mode in range(5):
        outs = []
        for fn in (getattr(rs, "ref_spectral_binary" + suf), getattr(lib, "orc_spectral_binary" + suf)):
            y = np.zeros(n1 + n2 + 8, dtype)
            outs.append((fn(ck.fptr(y), ck.fptr(a), n1, ck.fptr(b), n2, mode, 1, 32768), y))
        assert outs[0][0] == outs[1][0] > 0
        assert ck.rel_rms(outs[1][1], outs[0][1]) < tol, (mode,)
        if mode == 0 and n1 + n2 > 2:
            # Linear correlation: lags 0..n1-1 then the negative lags -(n2-1)..-1 (arrange_correlate :490-495)
            full = np.correlate(a.astype(np.float64), b.astype(np.float64), "full")         # lags -(n2-1) .. n1-1
            truth = np.concatenate([full[n2 - 1:], full[:n2 - 1]])
            assert ck.rel_rms(outs[1][1][: n1 + n2 - 1], truth) < tol
    ai = rng.uniform(-1, 1, max(n1 // 2, 0)).astype(dtype)
    bi = rng.uniform(-1, 1, n2).astype(dtype)
    for op in (0, 1):
        for mode in (0, 3, 4):
            outs = []
            for fn in (getattr(rs, "ref_spectral_binary_complex" + suf), getattr(lib, "orc_spectral_binary_complex" + suf)):
                yr, yi = np.zeros(n1 + n2 + 8, dtype), np.zeros(n1 + n2 + 8, dtype)
                size = fn(ck.fptr(yr), ck.fptr(yi), ck.fptr(a), n1, ck.fptr(ai), len(ai), ck.fptr(b), n2, ck.fptr(bi), n2, mode, op, 32768)
                outs.append((size, np.concatenate([yr, yi])))
            assert outs[0][0] == outs[1][0] > 0
            assert ck.rel_rms(outs[1][1], outs[0][1]) < tol, (op, mode)
    if n1 + n2 > 2:
        za = a.astype(np.complex128)
        za[:len(ai)] += 1j * ai
        zb = b.astype(np.float64) + 1j * bi
        yr, yi = np.zeros(n1 + n2 + 8, dtype), np.zeros(n1 + n2 + 8, dtype)
        getattr(lib, "orc_spectral_binary_complex" + suf)(ck.fptr(yr), ck.fptr(yi), ck.fptr(a), n1, ck.fptr(ai), len(ai), ck.fptr(b), n2, ck.fptr(bi), n2, 0, 0, 32768)
        truth = np.convolve(za, zb)
        assert ck.rel_rms(np.concatenate([yr[: n1 + n2 - 1], yi[: n1 + n2 - 1]]), np.concatenate([truth.real, truth.imag])) < tol


def test_spectral_convolve_fft_too_large_is_noop():
    rs, lib = ck.ref_spectral(), ck.oracle()
    a = np.ones(600, np.float64)
    for fn in (rs.ref_spectral_convolve_f64, lib.orc_spectral_convolve_f64):
        y = np.full(1300, 9.0)
        assert fn(ck.fptr(y), ck.fptr(a), 600, ck.fptr(a), 600, 0, 1024) == 0
        assert np.all(y == 9.0)


def test_reference_matrix_equals_sum_of_monos():
    """the uniform N x M reference used for configs 3/4 really is sum_i conv_i (NToMonoConvolve.cpp:35-43)."""
    rl = ck.ref()
    n_in, n_out, L, fft = 3, 2, 1500, 256
    m = rl.ref_matrix_create(n_in, n_out, L, fft, 0)
    irs = {}
    for o in range(n_out):
        for i in range(n_in):
            irs[(i, o)] = ck.synth_ir(L, o * n_in + i)
            assert rl.ref_matrix_set(m, i, o, ck.fptr(irs[(i, o)]), L) == 0
    n = 128 * 40
    x = np.stack([ck.synth_audio(n, i) for i in range(n_in)])
    y = np.zeros((n_out, n), np.float32)
    ins, outs = ck.planar_ptrs(x), ck.planar_ptrs(y)
    rl.ref_matrix_process(m, ins, outs, n)
    rl.ref_matrix_destroy(m)
    for o in range(n_out):
        truth = sum(ck.direct_convolve_delayed(irs[(i, o)], x[i], 128) for i in range(n_in))
        assert ck.rel_rms(y[o], truth) < 2e-6
