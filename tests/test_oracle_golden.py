"""Pins oracle/ against the committed golden vectors (tests/golden/golden.npz, produced from the
unmodified reference by tests/golden/make_golden.py).  CPU only; needs neither the reference nor a GPU.
"""
import os

import numpy as np
import pytest

import checkers as ck

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz"))
TOL = {"_f32": 2e-6, "_f64": 1e-13}


@pytest.mark.parametrize("suf", ["_f32", "_f64"])
@pytest.mark.parametrize("log2n", [1, 2, 3, 4, 5, 6, 9, 12])
@pytest.mark.parametrize("op", ["fft", "ifft", "rfft", "rifft"])
def test_fft_family(suf, log2n, op):
    key = "%s%s_%d" % (op, suf, log2n)
    lib = ck.oracle()
    s = getattr(lib, "orc_fft_setup_create" + suf)(13)
    re, im = (np.ascontiguousarray(a) for a in G[key + "_in"])
    getattr(lib, "orc_%s%s" % (op, suf))(s, ck.fptr(re), ck.fptr(im), log2n)
    getattr(lib, "orc_fft_setup_destroy" + suf)(s)
    assert ck.rel_rms(np.stack([re, im]), G[key + "_out"]) < TOL[suf]


@pytest.mark.parametrize("suf", ["_f32", "_f64"])
@pytest.mark.parametrize("log2n,in_length", [(5, 17), (8, 255), (10, 1024), (10, 700)])
def test_rfft_real(suf, log2n, in_length):
    key = "rfft_real%s_%d_%d" % (suf, log2n, in_length)
    lib = ck.oracle()
    x = np.ascontiguousarray(G[key + "_in"])
    n = 1 << log2n
    s = getattr(lib, "orc_fft_setup_create" + suf)(log2n)
    re, im, back = np.zeros(n >> 1, x.dtype), np.zeros(n >> 1, x.dtype), np.zeros(n, x.dtype)
    getattr(lib, "orc_rfft_real" + suf)(s, ck.fptr(x), ck.fptr(re), ck.fptr(im), in_length, log2n)
    assert ck.rel_rms(np.stack([re, im]), G[key + "_out"]) < TOL[suf]
    getattr(lib, "orc_rifft_real" + suf)(s, ck.fptr(re), ck.fptr(im), ck.fptr(back), log2n)
    getattr(lib, "orc_fft_setup_destroy" + suf)(s)
    assert ck.rel_rms(back, G[key + "_back"]) < TOL[suf]


@pytest.mark.parametrize("log2n", range(1, 18))
def test_zip_unzip_known_answer(log2n):
    """integer ramps must de-interleave/interleave exactly ("- Test/FFT_Tester/FFT_Tester/main.cpp":201-250)."""
    lib = ck.oracle()
    for dtype, suf in ((np.float32, "_f32"), (np.float64, "_f64")):
        n = 1 << log2n
        ramp = np.arange(n).astype(dtype)
        re, im, back = np.zeros(n // 2, dtype), np.zeros(n // 2, dtype), np.zeros(n, dtype)
        getattr(lib, "orc_unzip" + suf)(ck.fptr(ramp), ck.fptr(re), ck.fptr(im), log2n)
        assert np.array_equal(re, ramp[0::2]) and np.array_equal(im, ramp[1::2])
        getattr(lib, "orc_zip" + suf)(ck.fptr(re), ck.fptr(im), ck.fptr(back), log2n)
        assert np.array_equal(back, ramp)


@pytest.mark.parametrize("name", ["c1", "ragged", "phase", "slice", "trunc", "min"])
def test_pconv(name):
    fft, block, max_len, offset, length, reset_offset, err = (int(v) for v in G["pconv_%s_meta" % name])
    y, e = ck.oracle_pconv_run(fft, G["pconv_%s_ir" % name], G["pconv_%s_x" % name], block,
                               None if max_len < 0 else max_len, offset, length, reset_offset)
    assert e == err
    assert ck.rel_rms(y, G["pconv_%s_y" % name]) < 2e-6


def test_pconv_double():
    y, _ = ck.oracle_pconv_run(512, G["pconv64_ir"], G["pconv64_x"], 256, dtype=np.float64)
    assert ck.rel_rms(y, G["pconv64_y"]) < 1e-13


@pytest.mark.parametrize("mode,zero,sizes", [(0, 1, (256, 1024, 4096, 16384)), (1, 0, (256, 1024, 4096, 16384)),
                                             (2, 0, (1024, 4096, 16384, 0))])
def test_mono_latency_modes(mode, zero, sizes):
    lib = ck.oracle()
    ir, x = np.ascontiguousarray(G["mono_ir"]), np.ascontiguousarray(G["mono_x"])
    h = lib.orc_mono_create_f32(len(ir), zero, *sizes)
    assert lib.orc_mono_set_f32(h, ck.fptr(ir), len(ir), 1) == 0
    y = np.zeros_like(x)
    for pos in range(0, len(x), 512):
        n = min(512, len(x) - pos)
        lib.orc_mono_process_f32(h, ck.fptr(x[pos:]), ck.fptr(y[pos:]), n, 0)
    lib.orc_mono_destroy_f32(h)
    assert ck.rel_rms(y, G["mono_y_mode%d" % mode]) < 2e-6


def _oracle_matrix(irs, xs, sizes, zero=0):
    """sum over inputs of oracle MonoConvolves (NToMonoConvolve.cpp:35-43)."""
    lib = ck.oracle()
    n_out, n_in, L = irs.shape
    ys = np.zeros((n_out, xs.shape[1]), np.float32)
    for o in range(n_out):
        for i in range(n_in):
            ir = np.ascontiguousarray(irs[o, i])
            x = np.ascontiguousarray(xs[i])
            h = lib.orc_mono_create_f32(L, zero, *sizes)
            lib.orc_mono_set_f32(h, ck.fptr(ir), L, 1)
            lib.orc_mono_process_f32(h, ck.fptr(x), ck.fptr(ys[o]), len(x), 1)
            lib.orc_mono_destroy_f32(h)
    return ys


def test_convolver_shipped_mode():
    ys = _oracle_matrix(G["conv_irs"], G["conv_x"], (256, 1024, 4096, 16384))
    for o in range(ys.shape[0]):
        assert ck.rel_rms(ys[o], G["conv_y"][o]) < 2e-6


def test_uniform_matrix():
    fft = int(G["matrix_meta"][0])
    ys = _oracle_matrix(G["matrix_irs"], G["matrix_x"], (fft, 0, 0, 0))
    for o in range(ys.shape[0]):
        assert ck.rel_rms(ys[o], G["matrix_y"][o]) < 2e-6


@pytest.mark.parametrize("suf", ["_f32", "_f64"])
@pytest.mark.parametrize("n1,n2", [(1000, 300), (300, 1000), (64, 64), (7, 2), (1, 9)])
@pytest.mark.parametrize("mode", range(5))
def test_spectral_convolve(suf, n1, n2, mode):
    lib = ck.oracle()
    a = np.ascontiguousarray(G["spec%s_%d_%d_a" % (suf, n1, n2)])
    b = np.ascontiguousarray(G["spec%s_%d_%d_b" % (suf, n1, n2)])
    want = G["spec%s_%d_%d_m%d" % (suf, n1, n2, mode)]
    y = np.zeros(n1 + n2 + 8, a.dtype)
    size = getattr(lib, "orc_spectral_convolve" + suf)(ck.fptr(y), ck.fptr(a), n1, ck.fptr(b), n2, mode, 32768)
    assert size == len(want)
    assert ck.rel_rms(y[:size], want) < TOL[suf]


@pytest.mark.parametrize("suf", ["_f32", "_f64"])
@pytest.mark.parametrize("n1,n2", [(1000, 300), (300, 1000), (64, 64), (7, 2), (1, 9)])
def test_spectral_correlate_and_complex(suf, n1, n2):
    """oracle restatement of correlate (real, five edge modes) and of the complex-input convolve / correlate
    (Linear, Fold, FoldRepeat) against fixtures made by the unmodified reference."""
    lib = ck.oracle()
    a = np.ascontiguousarray(G["spec%s_%d_%d_a" % (suf, n1, n2)])
    b = np.ascontiguousarray(G["spec%s_%d_%d_b" % (suf, n1, n2)])
    for mode in range(5):
        want = G["corr%s_%d_%d_m%d" % (suf, n1, n2, mode)]
        y = np.zeros(n1 + n2 + 8, a.dtype)
        size = getattr(lib, "orc_spectral_binary" + suf)(ck.fptr(y), ck.fptr(a), n1, ck.fptr(b), n2, mode, 1, 32768)
        assert size == len(want)
        assert ck.rel_rms(y[:size], want) < TOL[suf], (mode,)
    ai = np.ascontiguousarray(G["cspec%s_%d_%d_ai" % (suf, n1, n2)])
    bi = np.ascontiguousarray(G["cspec%s_%d_%d_bi" % (suf, n1, n2)])
    for op in (0, 1):
        for mode in (0, 3, 4):
            want = G["cspec%s_%d_%d_op%d_m%d" % (suf, n1, n2, op, mode)]
            yr, yi = np.zeros(n1 + n2 + 8, a.dtype), np.zeros(n1 + n2 + 8, a.dtype)
            size = getattr(lib, "orc_spectral_binary_complex" + suf)(ck.fptr(yr), ck.fptr(yi), ck.fptr(a), n1, ck.fptr(ai), len(ai),
                                                                     ck.fptr(b), n2, ck.fptr(bi), n2, mode, op, 32768)
            assert size == want.shape[1]
            assert ck.rel_rms(np.concatenate([yr[:size], yi[:size]]), want.ravel()) < TOL[suf], (op, mode)


PHASES = ((0.0, 1.0), (0.3, 1.0), (0.5, 1.0), (1.0, 1.0), (0.8, 2.0))


@pytest.mark.parametrize("suf", ["_f32", "_f64"])
@pytest.mark.parametrize("size", [2, 5, 300, 1024])
def test_spectral_change_phase(suf, size):
    """oracle restatement of change_phase (minimum-phase cepstrum, interpolated, linear and maximum phase) against
    fixtures made by the unmodified reference."""
    lib = ck.oracle()
    x = np.ascontiguousarray(G["phase%s_%d_x" % (suf, size)])
    for k, (phase, tm) in enumerate(PHASES):
        want = G["phase%s_%d_k%d" % (suf, size, k)]
        y = np.zeros(4 * size + 16, x.dtype)
        n = getattr(lib, "orc_spectral_change_phase" + suf)(ck.fptr(y), ck.fptr(x), size, phase, tm)
        assert n == len(want)
        assert ck.rel_rms(y[:n], want) < (3e-6 if suf == "_f32" else 1e-12), (k,)
