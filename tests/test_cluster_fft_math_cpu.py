"""CPU restatement of the index arithmetic of the cluster transforms (hisstools_library_b200/csrc/hb_conv_cluster.cuh):
decimation in time over 8 ranks, the mirror-symmetric column sets of the forward split pass, the rank / index of the
mirror bin in the inverse, and the first-half-only cross pass.  numpy stands in for the local transforms; what is checked
is that the decomposition and the pairings reproduce the packed 2*DFT convention of HISSTools_FFT (reference:
HISSTools_FFT_Core.h:934-988, 1341-1374) exactly as the one-CTA kernels do."""
import numpy as np
import pytest

CS = 8


def packed_rfft(x):
    """forward real FFT in the reference's convention: 2 * DFT, Nyquist packed into the imaginary part of bin 0."""
    X = 2.0 * np.fft.fft(x)[: len(x) // 2 + 1]
    out = X[:-1].copy()
    out[0] = complex(X[0].real, X[-1].real)
    return out


def cluster_forward(frame):
    """k_fwd_cl, thread by thread (T threads per rank), returns the packed spectrum."""
    N = len(frame)
    M = N // 2
    L = M // CS
    T = L // 8
    h = T // 2
    z = frame[0::2] + 1j * frame[1::2]
    Y = [np.fft.fft(z[r::CS]) for r in range(CS)]                      # local transforms of z[8 n + r]
    out = np.zeros(M, complex)
    written = np.zeros(M, int)
    for r in range(CS):
        xs = np.zeros((8, T), complex)
        k2s = np.zeros(T, int)
        for t in range(T):
            a0 = r * h + (t if t < h else t - h)
            k2 = a0 if t < h else (L - a0 if a0 else L // 2)
            k2s[t] = k2
            v = np.array([Y[q][k2] * np.exp(-2j * np.pi * q * k2 / M) for q in range(8)])
            xs[:, t] = np.fft.fft(v)                                   # v[k1] = Z[k2 + L k1]
        for t in range(T):
            k2 = k2s[t]
            col0, colh = (r == 0 and t == 0), (r == 0 and t == h)
            pslot = t if (col0 or colh) else (t + h if t < h else t - h)
            for k1 in range(5 if col0 else 4):
                pk = (8 - k1) & 7 if col0 else 7 - k1
                k = k2 + L * k1
                q = (M - k) & (M - 1)
                a, b = xs[k1, t], xs[pk, pslot]
                if k == 0:
                    out[0] = complex(2 * (a.real + a.imag), 2 * (a.real - a.imag))
                    written[0] += 1
                    continue
                w = np.exp(-2j * np.pi * k / N)
                sr, si, dr, di = a.real + b.real, a.imag + b.imag, a.real - b.real, a.imag - b.imag
                u = w.real * si + w.imag * dr
                v_ = w.imag * si - w.real * dr
                out[k] = complex(sr + u, v_ + di)
                written[k] += 1
                if q != k:
                    out[q] = complex(sr - u, v_ - di)
                    written[q] += 1
    assert (written == 1).all()                                        # every bin is produced exactly once
    return out


def cluster_inverse_first_half(spec):
    """k_inv_cl: packed spectrum (Nyquist in imag of bin 0) -> first N/2 samples of the unscaled inverse real FFT."""
    M = len(spec)
    N = 2 * M
    L = M // CS
    T = L // 8
    nyq = spec[0].imag
    S = spec.copy()
    S[0] = complex(spec[0].real, 0.0)
    raw = [S[r::CS].copy() for r in range(CS)]                         # rank r holds bins 8 n + r
    loc = []
    for r in range(CS):
        s = np.zeros(L, complex)
        pr = (CS - r) & (CS - 1)
        for n in range(L):
            k = CS * n + r
            a = raw[r][n]
            b = raw[pr][L - 1 - n if r else (L - n) & (L - 1)]
            assert k == 0 or np.isclose(b, S[M - k])                   # the partner really is bin M - k
            if k == 0:
                zc = complex(a.real + nyq, a.real - nyq)
            else:
                w = np.exp(-2j * np.pi * k / N)
                sr, si, dr, di = a.real + b.real, a.imag + b.imag, a.real - b.real, a.imag - b.imag
                wr, wi = -w.real, w.imag
                u = wr * si + wi * dr
                v_ = wi * si - wr * dr
                zc = complex(sr + u, v_ + di)
            s[n] = complex(zc.imag, zc.real)                           # planes exchanged
        loc.append(np.fft.fft(s))
    y = np.zeros(N // 2)
    for r in range(CS):
        for t in range(T):
            k2 = r * T + t
            v = np.fft.fft(np.array([loc[q][k2] * np.exp(-2j * np.pi * q * k2 / M) for q in range(8)]))
            for k1 in range(4):
                kk = k2 + L * k1
                y[2 * kk], y[2 * kk + 1] = v[k1].imag, v[k1].real
    return y


@pytest.mark.parametrize("log2n", [12, 13, 14])
def test_cluster_forward_matches_packed_rfft(log2n):
    rng = np.random.default_rng(log2n)
    x = rng.standard_normal(1 << log2n)
    got = cluster_forward(x)
    want = packed_rfft(x)
    assert np.max(np.abs(got - want)) <= 1e-9 * np.max(np.abs(want))


@pytest.mark.parametrize("log2n", [12, 13, 14])
def test_cluster_inverse_first_half_round_trip(log2n):
    """rfft -> rifft of the reference's pair is 2N times the input; the cluster inverse keeps the first half."""
    rng = np.random.default_rng(100 + log2n)
    N = 1 << log2n
    x = rng.standard_normal(N)
    y = cluster_inverse_first_half(packed_rfft(x))
    assert np.max(np.abs(y - 2 * N * x[: N // 2])) <= 1e-9 * 2 * N
