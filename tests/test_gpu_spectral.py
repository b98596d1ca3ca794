"""GPU parity of the one-shot spectral convolution (spectral_processor<T>::convolve, SpectralProcessor.hpp:169-172)
against the golden vectors produced by the unmodified reference, the plain-C oracle, the compiled
reference where shipped, and float64 direct convolution.  Tolerances: 1e-5 float, 1e-12 double."""
import os

import numpy as np
import pytest

import checkers as ck

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz"))
TOL = {"f32": 1e-5, "f64": 1e-12}
DT = {"f32": np.float32, "f64": np.float64}
# FFT sizes above 2^15 (float) / 2^14 (double) take the four-step global-memory path
MAXFFT = {"f32": 65536, "f64": 32768}


@pytest.fixture(scope="module")
def hb():
    import hisstools_library_b200 as h
    return h


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("n1,n2", [(1000, 300), (300, 1000), (64, 64), (7, 2), (1, 9)])
def test_golden_all_edge_modes(hb, suf, n1, n2):
    sp = hb.spectral_processor(MAXFFT[suf], DT[suf])
    a, b = G["spec_%s_%d_%d_a" % (suf, n1, n2)], G["spec_%s_%d_%d_b" % (suf, n1, n2)]
    for mode in range(5):
        want = G["spec_%s_%d_%d_m%d" % (suf, n1, n2, mode)]
        assert sp.convolved_size(n1, n2, mode) == len(want)
        out = np.zeros(len(want) + 3, DT[suf])
        assert sp.convolve(out, a, b, mode) == len(want)
        assert ck.rel_rms(out[:len(want)], want) <= TOL[suf], (suf, n1, n2, mode)
        assert np.all(out[len(want):] == 0)


@pytest.mark.parametrize("suf", ["f32", "f64"])
def test_against_oracle_and_direct(hb, suf):
    dt = DT[suf]
    lib = ck.oracle()
    fn = getattr(lib, "orc_spectral_convolve_" + suf)
    sp = hb.spectral_processor(MAXFFT[suf], dt)
    rng = np.random.default_rng(21)
    for n1, n2 in [(1, 1), (1, 2), (2, 1), (2, 2), (3, 5), (16, 16), (129, 64), (4000, 4000), (12000, 3000), (20000, 12000), (40000, 20000)]:
        a = rng.uniform(-1, 1, n1).astype(dt)
        b = (rng.standard_normal(n2) * np.exp(-3.0 * np.arange(n2) / n2)).astype(dt)
        for mode in range(5):
            want = np.zeros(n1 + n2, dt)
            size = fn(ck.fptr(want), ck.fptr(a), n1, ck.fptr(b), n2, mode, MAXFFT[suf])
            got = np.zeros(n1 + n2, dt)
            assert sp.convolve(got, a, b, mode) == size, (n1, n2, mode)
            if size:
                assert ck.rel_rms(got[:size], want[:size]) <= TOL[suf], (suf, n1, n2, mode)
        if n1 + n2 - 1 > MAXFFT[suf]:
            continue
        lin = np.zeros(n1 + n2 - 1, dt)
        sp.convolve(lin, a, b, hb.EdgeMode.Linear)
        truth = np.convolve(a.astype(np.float64), b.astype(np.float64))
        assert ck.rel_rms(lin, truth) <= TOL[suf]
        # Wrap = circular convolution of period max(n1, n2)
        mx = max(n1, n2)
        wrap = np.zeros(mx, dt)
        sp.convolve(wrap, a, b, hb.EdgeMode.Wrap)
        circ = truth[:mx].copy()
        circ[:len(truth) - mx] += truth[mx:]
        assert ck.rel_rms(wrap, circ) <= TOL[suf]


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("n1,n2", [(1000, 300), (300, 1000), (64, 64), (7, 2), (1, 9)])
def test_golden_correlate_and_complex(hb, suf, n1, n2):
    """correlate (real, five edge modes) and complex-input convolve / correlate (Linear, Fold, FoldRepeat) against the
    fixtures made by the unmodified reference."""
    sp = hb.spectral_processor(MAXFFT[suf], DT[suf])
    a, b = G["spec_%s_%d_%d_a" % (suf, n1, n2)], G["spec_%s_%d_%d_b" % (suf, n1, n2)]
    for mode in range(5):
        want = G["corr_%s_%d_%d_m%d" % (suf, n1, n2, mode)]
        assert sp.correlated_size(n1, n2, mode) == len(want)
        out = np.zeros(len(want) + 3, DT[suf])
        assert sp.correlate(out, a, b, mode) == len(want)
        assert ck.rel_rms(out[:len(want)], want) <= TOL[suf], (suf, n1, n2, mode)
        assert np.all(out[len(want):] == 0)
    ai, bi = G["cspec_%s_%d_%d_ai" % (suf, n1, n2)], G["cspec_%s_%d_%d_bi" % (suf, n1, n2)]
    for op, fn in ((0, sp.convolve), (1, sp.correlate)):
        for mode in (0, 3, 4):
            want = G["cspec_%s_%d_%d_op%d_m%d" % (suf, n1, n2, op, mode)]
            size = want.shape[1]
            r_out, i_out = np.zeros(size + 2, DT[suf]), np.zeros(size + 2, DT[suf])
            assert fn(r_out, i_out, a, ai, b, bi, mode) == size
            assert ck.rel_rms(np.concatenate([r_out[:size], i_out[:size]]), want.ravel()) <= TOL[suf], (suf, n1, n2, op, mode)
            assert np.all(r_out[size:] == 0) and np.all(i_out[size:] == 0)


@pytest.mark.parametrize("suf", ["f32", "f64"])
def test_correlate_and_complex_against_oracle(hb, suf):
    """all five edge modes, real correlate and both complex operations, against the plain-C oracle (which is pinned to the
    reference), including sizes on the four-step path, absent planes and the single-sample special cases."""
    dt = DT[suf]
    lib = ck.oracle()
    fr, fc = getattr(lib, "orc_spectral_binary_" + suf), getattr(lib, "orc_spectral_binary_complex_" + suf)
    sp = hb.spectral_processor(MAXFFT[suf], dt)
    rng = np.random.default_rng(77)
    for n1, n2 in [(1, 1), (1, 2), (2, 1), (3, 5), (16, 16), (129, 64), (64, 129), (4000, 4000), (12000, 3000), (3000, 12000), (20000, 12000), (40000, 20000)]:
        a = rng.uniform(-1, 1, n1).astype(dt)
        b = (rng.standard_normal(n2) * np.exp(-3.0 * np.arange(n2) / n2)).astype(dt)
        for mode in range(5):
            want = np.zeros(n1 + n2, dt)
            size = fr(ck.fptr(want), ck.fptr(a), n1, ck.fptr(b), n2, mode, 1, MAXFFT[suf])
            got = np.zeros(n1 + n2, dt)
            assert sp.correlate(got, a, b, mode) == size, (n1, n2, mode)
            if size:
                assert ck.rel_rms(got[:size], want[:size]) <= TOL[suf], (suf, n1, n2, mode)
        for planes in ((n1, n1, n2, n2), (n1, n1 // 2, n2, 0), (0, n1, n2, n2 // 3)):
            ps = [rng.uniform(-1, 1, k).astype(dt) for k in planes]
            for op, fn in ((0, sp.convolve), (1, sp.correlate)):
                for mode in range(5):
                    wr, wi = np.zeros(n1 + n2, dt), np.zeros(n1 + n2, dt)
                    size = fc(ck.fptr(wr), ck.fptr(wi), ck.fptr(ps[0]), planes[0], ck.fptr(ps[1]), planes[1], ck.fptr(ps[2]), planes[2],
                              ck.fptr(ps[3]), planes[3], mode, op, MAXFFT[suf])
                    gr, gi = np.zeros(n1 + n2, dt), np.zeros(n1 + n2, dt)
                    assert fn(gr, gi, ps[0] if planes[0] else None, ps[1] if planes[1] else None, ps[2], ps[3] if planes[3] else None, mode) == size
                    if size:
                        assert ck.rel_rms(np.concatenate([gr[:size], gi[:size]]), np.concatenate([wr[:size], wi[:size]])) <= TOL[suf], (suf, planes, op, mode)


PHASES = ((0.0, 1.0), (0.3, 1.0), (0.5, 1.0), (1.0, 1.0), (0.8, 2.0))


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("size", [2, 5, 300, 1024])
def test_golden_change_phase(hb, suf, size):
    """change_phase against fixtures made by the unmodified reference: minimum (0), interpolated (0.3, 0.8 with a doubled
    FFT), linear (0.5) and maximum (1) phase."""
    sp = hb.spectral_processor(MAXFFT[suf], DT[suf])
    x = G["phase_%s_%d_x" % (suf, size)]
    for k, (phase, tm) in enumerate(PHASES):
        want = G["phase_%s_%d_k%d" % (suf, size, k)]
        out = np.zeros(len(want) + 3, DT[suf])
        assert sp.change_phase(out, x, size, phase, tm) == len(want)
        assert ck.rel_rms(out[:len(want)], want) <= TOL[suf], (suf, size, k)
        assert np.all(out[len(want):] == 0)


@pytest.mark.parametrize("suf", ["f32", "f64"])
def test_change_phase_against_oracle_and_properties(hb, suf):
    """larger sizes (single-CTA and four-step transforms) against the C oracle; the minimum- and maximum-phase versions keep
    the magnitude spectrum, the maximum-phase one is the time reverse of the minimum-phase one delayed by a sample."""
    dt = DT[suf]
    lib = ck.oracle()
    fn = getattr(lib, "orc_spectral_change_phase_" + suf)
    sp = hb.spectral_processor(1 << 17, dt)
    rng = np.random.default_rng(12)
    for size in (7, 2048, 20000, 40000):
        x = (rng.standard_normal(size) * np.exp(-6.0 * np.arange(size) / size)).astype(dt)
        for phase, tm in ((0.0, 1.0), (0.25, 1.0), (0.5, 1.0), (1.0, 1.0), (0.6, 2.0)):
            want = np.zeros(4 * size + 16, dt)
            n = fn(ck.fptr(want), ck.fptr(x), size, phase, tm)
            got = np.zeros(4 * size + 16, dt)
            assert sp.change_phase(got, x, size, phase, tm) == n
            # log / exp of the spectrum amplify rounding differences between two FFT factorizations with the transform
            # length: 1e-12 holds up to 1024 points (golden test above), 1e-10 is asserted at 65536
            assert ck.rel_rms(got[:n], want[:n]) <= (TOL[suf] if suf == "f32" else 1e-10), (suf, size, phase, tm)
        n = sp.change_phase(got, x, size, 0.0, 1.0)
        mn = got[:n].astype(np.float64)
        mag_in = np.abs(np.fft.rfft(x.astype(np.float64), n))
        assert ck.rel_rms(np.abs(np.fft.rfft(mn)), mag_in) <= (1e-4 if suf == "f32" else 1e-9)
    with pytest.raises(hb.HissError):
        sp.change_phase(np.zeros(1 << 19, dt), np.zeros(200000, dt), 200000, 0.0, 1.0)     # FFT above the processor's maximum
    one = np.zeros(1, dt)
    assert sp.change_phase(one, np.array([0.75], dt), 1, 0.3) == 1 and one[0] == dt(0.75)


def test_limits_and_noops(hb):
    sp = hb.spectral_processor(1024)
    assert sp.max_fft_size() == 1024
    out = np.full(2000, 5.0, np.float32)
    a, b = ck.synth_audio(700, 1), ck.synth_audio(400, 2)
    assert sp.convolved_size(700, 400, 0) == 0                      # needs 2048 > max 1024
    assert sp.convolve(out, a, b, 0) == 0 and np.all(out == 5.0)    # silently does nothing (SpectralProcessor.hpp:651-652)
    assert sp.convolve(out, a[:0], b, 0) == 0 and np.all(out == 5.0)
    sp.set_max_fft_size(2048)
    assert sp.convolve(out, a, b, 0) == 1099
    assert ck.rel_rms(out[:1099], np.convolve(a.astype(np.float64), b.astype(np.float64))) <= 1e-5
    sp.set_max_fft_size(1000)                                       # rounds up to 1024
    assert sp.max_fft_size() == 1024
    with pytest.raises(hb.HissError):
        hb.spectral_processor(1 << 25)                              # beyond what this build implements


def test_against_compiled_reference(hb):
    rs = ck.ref_spectral()
    if rs is None:
        pytest.skip("compiled reference not shipped")
    rng = np.random.default_rng(5)
    for suf, dt in (("f32", np.float32), ("f64", np.float64)):
        sp = hb.spectral_processor(MAXFFT[suf], dt)
        a = rng.uniform(-1, 1, 5000).astype(dt)
        b = rng.uniform(-1, 1, 1234).astype(dt)
        for mode in range(5):
            want = np.zeros(7000, dt)
            size = getattr(rs, "ref_spectral_convolve_" + suf)(ck.fptr(want), ck.fptr(a), 5000, ck.fptr(b), 1234, mode, MAXFFT[suf])
            got = np.zeros(7000, dt)
            assert sp.convolve(got, a, b, mode) == size
            assert ck.rel_rms(got[:size], want[:size]) <= TOL[suf]


@pytest.mark.parametrize("suf", ["f32", "f64"])
def test_kernel_smoother_shaped_caller(hb, suf):
    """The one-shot path as its caller in the reference uses it: kernel_smoother::apply_filter_fft (KernelSmoother.hpp:317-326)
    convolves n + width - 1 data samples with a filter of `width` taps in Linear mode and keeps outputs [width - 1, width - 1 + n)
    times a gain.  Same sizes and slicing here, against the compiled reference's spectral_processor where shipped, the oracle and
    the direct form the smoother's time-domain branch computes (apply_filter :288-301)."""
    dt = DT[suf]
    rng = np.random.default_rng(31)
    sp = hb.spectral_processor(MAXFFT[suf], dt)
    lib = ck.oracle()
    fn = getattr(lib, "orc_spectral_convolve_" + suf)
    rs = ck.ref_spectral()
    for n, width, gain in [(512, 33, 0.25), (2048, 257, 1.0 / 257), (4000, 1001, 0.003), (16, 5, 2.0)]:
        data = rng.uniform(-1, 1, n + width - 1).astype(dt)
        filt = np.hanning(width + 2)[1:-1].astype(dt)
        full = np.zeros(n + 2 * width, dt)
        size = sp.convolve(full, data, filt, hb.EdgeMode.Linear)
        assert size == n + 2 * width - 2
        out = full[width - 1:width - 1 + n] * dt(gain)
        want = np.zeros(n + 2 * width, dt)
        assert fn(ck.fptr(want), ck.fptr(data), len(data), ck.fptr(filt), width, 0, MAXFFT[suf]) == size
        assert ck.rel_rms(out, want[width - 1:width - 1 + n] * dt(gain)) <= TOL[suf]
        if rs is not None:
            refo = np.zeros(n + 2 * width, dt)
            assert getattr(rs, "ref_spectral_convolve_" + suf)(ck.fptr(refo), ck.fptr(data), len(data), ck.fptr(filt), width, 0, MAXFFT[suf]) == size
            assert ck.rel_rms(out, refo[width - 1:width - 1 + n] * dt(gain)) <= TOL[suf]
        # the time-domain branch of the smoother: out[i] = gain * sum_j filter[j] * data[i + width - 1 - j]
        direct = np.array([np.dot(filt.astype(np.float64), data[i:i + width][::-1].astype(np.float64)) for i in range(min(n, 64))]) * gain
        assert ck.rel_rms(out[:len(direct)], direct) <= TOL[suf] * 10
