"""GPU parity of the hisstools_fft family (CUDA path through the C ABI) against the CPU oracle and the
golden vectors produced by the unmodified reference.  Shaped after the reference's FFT_Tester
("- Test/FFT_Tester/FFT_Tester/main.cpp": every log2n, fft/ifft/rfft/rifft, float and double, inputs
uniform[-1,1]) plus the numeric checks the reference never had.
Tolerances (BASELINE.json north_star): <= 1e-5 relative RMS float, <= 1e-12 double.
"""
import os

import numpy as np
import pytest

import checkers as ck

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz"))
TOL = {"_f32": 1e-5, "_f64": 1e-12}
DT = {"_f32": np.float32, "_f64": np.float64}
# largest complex transform exercised here (the reference's FFT_Tester runs every log2n up to 21): sizes above
# 2^14 float / 2^13 double leave the single-CTA shared-memory path for the four-step global-memory path
MAX_C = {"_f32": 20, "_f64": 20}


@pytest.fixture(scope="module")
def hb():
    import hisstools_library_b200 as h
    return h


@pytest.fixture(scope="module")
def setups(hb):
    s = {"_f32": hb.hisstools_create_setup(21, np.float32), "_f64": hb.hisstools_create_setup(21, np.float64)}
    yield s
    for v in s.values():
        hb.hisstools_destroy_setup(v)


def _oracle_op(op, suf, re, im, log2n):
    lib = ck.oracle()
    s = getattr(lib, "orc_fft_setup_create" + suf)(max(log2n, 4))
    getattr(lib, "orc_%s%s" % (op, suf))(s, ck.fptr(re), ck.fptr(im), log2n)
    getattr(lib, "orc_fft_setup_destroy" + suf)(s)


@pytest.mark.parametrize("suf", ["_f32", "_f64"])
@pytest.mark.parametrize("op", ["fft", "ifft", "rfft", "rifft"])
def test_every_size_against_oracle(hb, setups, suf, op):
    rng = np.random.default_rng(7)
    fn = getattr(hb, "hisstools_" + op)
    top = MAX_C[suf] + (1 if op in ("rfft", "rifft") else 0)
    for log2n in range(0, top + 1):
        n = 1 << log2n
        planes = n if op in ("fft", "ifft") else max(n >> 1, 1)
        re = rng.uniform(-1, 1, planes).astype(DT[suf])
        im = rng.uniform(-1, 1, planes).astype(DT[suf])
        want_re, want_im = re.copy(), im.copy()
        if not (op in ("rfft", "rifft") and log2n == 0):
            _oracle_op(op, suf, want_re, want_im, log2n)
        fn(setups[suf], hb.Split(re, im), log2n)
        err = ck.rel_rms(np.stack([re, im]), np.stack([want_re, want_im]))
        assert err <= TOL[suf], (op, suf, log2n, err)


@pytest.mark.parametrize("suf", ["_f32", "_f64"])
@pytest.mark.parametrize("log2n", [1, 2, 3, 4, 5, 6, 9, 12])
@pytest.mark.parametrize("op", ["fft", "ifft", "rfft", "rifft"])
def test_golden(hb, setups, suf, log2n, op):
    key = "%s%s_%d" % (op, suf, log2n)
    re, im = (np.ascontiguousarray(a) for a in G[key + "_in"])
    getattr(hb, "hisstools_" + op)(setups[suf], hb.Split(re, im), log2n)
    assert ck.rel_rms(np.stack([re, im]), G[key + "_out"]) <= TOL[suf]


@pytest.mark.parametrize("suf", ["_f32", "_f64"])
@pytest.mark.parametrize("log2n,in_length", [(5, 17), (8, 255), (10, 1024), (10, 700)])
def test_out_of_place_real_golden(hb, setups, suf, log2n, in_length):
    """zero padding, odd lengths (Core:1258-1287) and the zipped inverse (HISSTools_FFT.h:269,282)."""
    key = "rfft_real%s_%d_%d" % (suf, log2n, in_length)
    x = np.ascontiguousarray(G[key + "_in"])
    n = 1 << log2n
    sp = hb.Split.zeros(n >> 1, x.dtype)
    hb.hisstools_rfft(setups[suf], x, sp, in_length, log2n)
    assert ck.rel_rms(np.stack([sp.realp, sp.imagp]), G[key + "_out"]) <= TOL[suf]
    back = np.zeros(n, x.dtype)
    hb.hisstools_rifft(setups[suf], sp, back, log2n)
    assert ck.rel_rms(back, G[key + "_back"]) <= TOL[suf]
    # the planes are left holding the de-interleaved result, as in the reference
    assert np.array_equal(sp.realp, back[0::2]) and np.array_equal(sp.imagp, back[1::2])


def test_float_input_double_setup(hb, setups):
    """hisstools_rfft(FFT_SETUP_D, const float*, ...) (HISSTools_FFT.h:208)."""
    x = np.random.default_rng(3).uniform(-1, 1, 1000).astype(np.float32)
    sp = hb.Split.zeros(512, np.float64)
    hb.hisstools_rfft(setups["_f64"], x, sp, 1000, 10)
    want = hb.Split.zeros(512, np.float64)
    hb.hisstools_rfft(setups["_f64"], x.astype(np.float64), want, 1000, 10)
    assert np.array_equal(sp.realp, want.realp) and np.array_equal(sp.imagp, want.imagp)


@pytest.mark.parametrize("suf", ["_f32", "_f64"])
def test_conventions_and_round_trips(hb, setups, suf):
    """2*DFT with packed DC/Nyquist, rifft(rfft(x)) = 2N x, ifft(fft(z)) = N z."""
    dt = DT[suf]
    rng = np.random.default_rng(11)
    for log2n in (4, 7, 11, 13, 16, 19):
        n = 1 << log2n
        x = rng.uniform(-1, 1, n).astype(dt)
        sp = hb.Split.zeros(n >> 1, dt)
        hb.hisstools_unzip(x, sp, log2n)
        hb.hisstools_rfft(setups[suf], sp, log2n)
        spec = np.fft.rfft(x.astype(np.float64))
        got = sp.realp[1:] + 1j * sp.imagp[1:]
        tol = TOL[suf]
        assert ck.rel_rms(np.concatenate([got.real, got.imag]), np.concatenate([2 * spec[1:-1].real, 2 * spec[1:-1].imag])) <= tol
        assert abs(sp.realp[0] - 2 * spec[0].real) <= tol * max(1.0, abs(2 * spec[0].real)) * 10
        assert abs(sp.imagp[0] - 2 * spec[-1].real) <= tol * max(1.0, abs(2 * spec[-1].real)) * 10 + tol
        hb.hisstools_rifft(setups[suf], sp, log2n)
        back = np.zeros(n, dt)
        hb.hisstools_zip(sp, back, log2n)
        assert ck.rel_rms(back, 2.0 * n * x.astype(np.float64)) <= tol
        zr, zi = rng.uniform(-1, 1, n).astype(dt), rng.uniform(-1, 1, n).astype(dt)
        z = hb.Split(zr.copy(), zi.copy())
        hb.hisstools_fft(setups[suf], z, log2n)
        ref = np.fft.fft(zr.astype(np.float64) + 1j * zi.astype(np.float64))
        assert ck.rel_rms(np.concatenate([z.realp, z.imagp]), np.concatenate([ref.real, ref.imag])) <= tol
        hb.hisstools_ifft(setups[suf], z, log2n)
        assert ck.rel_rms(np.concatenate([z.realp, z.imagp]), n * np.concatenate([zr, zi]).astype(np.float64)) <= tol


def test_linearity_full_size(hb, setups):
    """size-independent property at the largest shared-memory size: F(a x + b y) = a F(x) + b F(y)."""
    rng = np.random.default_rng(5)
    n = 1 << 15
    x, y = rng.uniform(-1, 1, n).astype(np.float32), rng.uniform(-1, 1, n).astype(np.float32)
    out = []
    for sig in (x, y, (0.5 * x - 2.0 * y).astype(np.float32)):
        sp = hb.Split.zeros(n >> 1)
        hb.hisstools_rfft(setups["_f32"], sig, sp, n, 15)
        out.append(np.concatenate([sp.realp, sp.imagp]).astype(np.float64))
    assert ck.rel_rms(out[2], 0.5 * out[0] - 2.0 * out[1]) <= 1e-5


@pytest.mark.parametrize("suf", ["_f32", "_f64"])
def test_out_of_place_real_large(hb, setups, suf):
    """zero-padded out-of-place real transform and zipped inverse on the four-step path (2^17 points)."""
    dt = DT[suf]
    log2n, in_length = 17, 100001
    n = 1 << log2n
    x = np.random.default_rng(17).uniform(-1, 1, in_length).astype(dt)
    sp = hb.Split.zeros(n >> 1, dt)
    hb.hisstools_rfft(setups[suf], x, sp, in_length, log2n)
    xp = np.zeros(n)
    xp[:in_length] = x
    spec = 2 * np.fft.rfft(xp)
    want = np.concatenate([spec.real[:n >> 1], spec.imag[:n >> 1]])
    want[n >> 1] = spec.real[n >> 1]                         # Nyquist packed into imagp[0]
    assert ck.rel_rms(np.concatenate([sp.realp, sp.imagp]), want) <= TOL[suf]
    back = np.zeros(n, dt)
    hb.hisstools_rifft(setups[suf], sp, back, log2n)
    assert ck.rel_rms(back, 2.0 * n * xp) <= TOL[suf]
    assert np.array_equal(sp.realp, back[0::2]) and np.array_equal(sp.imagp, back[1::2])
