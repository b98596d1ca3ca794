"""The C++ drop-in boundary: tests/cpp/dropin_test.cpp is written against the reference's C++ API and
compiled against include/HISSTools_FFT + include/HIRT_Multichannel_Convolution, linked to
libhisstools_b200.so.  Its outputs are checked here against the oracle / the compiled reference /
float64 direct convolution on the same (LCG-generated) inputs.
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import checkers as ck

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL32 = 1e-5


def lcg_noise(n, seed):
    s = np.uint64(seed)
    out = np.empty(n, np.float32)
    state = int(seed)
    for k in range(n):
        state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
        out[k] = np.float32(state >> 8) * np.float32(1.0 / 8388608.0) - np.float32(1.0)
    return out


def decaying(n, seed):
    v = lcg_noise(n, seed)
    k = np.arange(n, dtype=np.float32)
    return (v * (np.float32(1.0) - k / np.float32(n))).astype(np.float32)


def build_program(extra=()):
    from hisstools_library_b200 import build
    lib = build.build()
    d = tempfile.mkdtemp(prefix="hb_dropin_")
    exe = os.path.join(d, "dropin_test")
    subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "dropin_test.cpp"),
                    "-L" + os.path.dirname(lib), "-lhisstools_b200", "-Wl,-rpath," + os.path.dirname(lib), "-o", exe] + list(extra), check=True)
    return exe, d


def test_cpp_headers_compile_against_reference_style_caller():
    """CPU-only: the caller builds against the drop-in headers (same class / function names as the reference)."""
    exe, _ = build_program()
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_cpp_dropin_results():
    _run_and_check(0)


@pytest.mark.gpu
def test_cpp_dropin_on_several_gpus():
    """The same caller with the device-list constructors of HISSTools::Convolver: one object, one process(ins, outs, ...) call
    with host pointers, the matrix dealt to every GPU of the box (section 10 of dropin_test.cpp)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    _run_and_check(8 if n >= 8 else (4 if n >= 4 else 2))


def _run_and_check(ndev):
    exe, d = build_program()
    out = os.path.join(d, "out.bin")
    wav = os.path.join(ROOT, "tests", "golden", "audio", "wave_i24.wav")
    res = subprocess.run([exe, out, wav] + ([str(ndev)] if ndev else []), capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "kernel launches" in res.stdout and int(res.stdout.split()[1]) > 0
    data = np.fromfile(out, np.float32)
    pos = [0]

    def take(n):
        v = data[pos[0]:pos[0] + n]
        pos[0] += n
        assert len(v) == n
        return v

    def code():
        return int(take(1)[0])

    lib = ck.oracle()
    # 1. FFT family
    x = lcg_noise(1024, 1)
    re, im = np.zeros(512, np.float32), np.zeros(512, np.float32)
    s = lib.orc_fft_setup_create_f32(10)
    lib.orc_rfft_real_f32(s, ck.fptr(x), ck.fptr(re), ck.fptr(im), 1000, 10)
    assert ck.rel_rms(np.stack([take(512), take(512)]), np.stack([re, im])) <= TOL32
    y = np.zeros(1024, np.float32)
    lib.orc_rifft_real_f32(s, ck.fptr(re), ck.fptr(im), ck.fptr(y), 10)
    got = take(1024)
    assert ck.rel_rms(got, y) <= TOL32
    xp = x.copy(); xp[1000:] = 0
    assert ck.rel_rms(got, 2048.0 * xp) <= TOL32                          # rifft(rfft(x)) = 2N x
    cr, ci = lcg_noise(256, 2), lcg_noise(256, 3)
    z = np.fft.fft(cr.astype(np.float64) + 1j * ci.astype(np.float64))
    assert ck.rel_rms(np.stack([take(256), take(256)]), np.stack([z.real, z.imag])) <= TOL32
    lib.orc_fft_setup_destroy_f32(s)
    xf = lcg_noise(256, 4).astype(np.float64)
    X = 2 * np.fft.rfft(xf)
    want_re, want_im = X.real[:128].copy(), X.imag[:128].copy()
    want_im[0] = X.real[128]
    assert ck.rel_rms(np.stack([take(128), take(128)]), np.stack([want_re, want_im])) <= 1e-6

    # 2. PartitionedConvolve
    x, ir = lcg_noise(6000, 10), decaying(3000, 11)
    assert code() == 0 and code() == 7                                    # no IR: false, output untouched
    assert code() == 0 and code() == 11
    want, _ = ck.oracle_pconv_run(512, ir, x, 256)
    assert ck.rel_rms(take(6000), want) <= TOL32

    # 3. MonoConvolve(5000, kLatencyShort): net delay 128
    x, ir = lcg_noise(4096, 20), decaying(5000, 21)
    assert code() == 0
    assert ck.rel_rms(take(4096), ck.direct_convolve_delayed(ir, x, 128)) <= TOL32
    assert code() == 1                                                    # invalid size order throws

    # 4. NToMonoConvolve(3, 2000, kLatencyMedium): delay 512
    xs = [lcg_noise(4096, 30 + i) for i in range(3)]
    irs = [decaying(2000, 40 + i) for i in range(3)]
    assert [code() for _ in range(4)] == [0, 0, 0, 1]
    truth = sum(ck.direct_convolve_delayed(irs[i], xs[i], 512) for i in range(3))
    assert ck.rel_rms(take(4096), truth) <= TOL32

    # 5. Convolver(3, 2, kLatencyZero): zero delay, IRs longer than the default allocation
    xs = [lcg_noise(8192, 50 + i) for i in range(3)]
    irs = [decaying(20000, 60 + p) for p in range(6)]
    assert code() == 4
    assert [code() for _ in range(6)] == [0] * 6
    assert code() == 2
    for o in range(2):
        truth = sum(ck.direct_convolve_delayed(irs[o * 3 + i], xs[i], 0) for i in range(3))
        assert ck.rel_rms(take(8192), truth) <= TOL32
    truth = sum(ck.direct_convolve_delayed(irs[i], xs[i][:1024], 0) for i in range(3))
    assert ck.rel_rms(take(1024), truth) <= TOL32                         # double I/O

    # 6. Convolver(4, kLatencyShort): parallel channels
    xs = [lcg_noise(2048, 70 + i) for i in range(4)]
    irs = [decaying(1500, 80 + i) for i in range(4)]
    assert [code() for _ in range(5)] == [0, 0, 0, 0, 1]
    for i in range(4):
        assert ck.rel_rms(take(2048), ck.direct_convolve_delayed(irs[i], xs[i], 128)) <= TOL32
    # 7. spectral_processor<float>::convolve
    a, b = lcg_noise(900, 90), decaying(250, 91)
    assert code() == 1149
    assert ck.rel_rms(take(1149), np.convolve(a.astype(np.float64), b.astype(np.float64))) <= TOL32
    want = np.zeros(1200, np.float32)
    size = lib.orc_spectral_convolve_f32(ck.fptr(want), ck.fptr(a), 900, ck.fptr(b), 250, 2, 32768)
    assert size == 900 and ck.rel_rms(take(900), want[:900]) <= TOL32
    # 8. correlate (Wrap), complex-input convolve (Linear), change_phase (minimum phase, FFT 1024)
    a, b, ai, bi = lcg_noise(700, 92), decaying(300, 93), lcg_noise(350, 94), lcg_noise(300, 95)
    assert code() == 700
    want = np.zeros(1100, np.float32)
    assert lib.orc_spectral_binary_f32(ck.fptr(want), ck.fptr(a), 700, ck.fptr(b), 300, 1, 1, 32768) == 700
    assert ck.rel_rms(take(700), want[:700]) <= TOL32
    wr, wi = np.zeros(1100, np.float32), np.zeros(1100, np.float32)
    assert lib.orc_spectral_binary_complex_f32(ck.fptr(wr), ck.fptr(wi), ck.fptr(a), 700, ck.fptr(ai), 350, ck.fptr(b), 300, ck.fptr(bi), 300, 0, 0, 32768) == 999
    assert ck.rel_rms(np.concatenate([take(999), take(999)]), np.concatenate([wr[:999], wi[:999]])) <= TOL32
    want = np.zeros(1100, np.float32)
    assert lib.orc_spectral_change_phase_f32(ck.fptr(want), ck.fptr(b), 300, 0.0, 3.0) == 1024
    assert ck.rel_rms(take(1024), want[:1024]) <= TOL32
    # 9. IAudioFile on a fixture written and read back by the reference (bit-exact)
    GA = np.load(os.path.join(ROOT, "tests", "golden", "golden_audio.npz"))
    meta = GA["wave_i24_wav_meta"]
    assert code() == 1 and code() == 0
    assert [code() for _ in range(4)] == [int(meta[1]), int(meta[2]), int(meta[5]), int(meta[6])]
    assert code() == int(GA["wave_i24_wav_rate"][0]) and code() == 24 and code() == 3 * int(meta[5])
    assert code() == 117
    assert np.array_equal(take(100).view(np.uint32), GA["wave_i24_wav_ch1_from17_f32"].view(np.uint32))
    assert np.array_equal(take(int(meta[5]) * int(meta[6])).view(np.uint32), GA["wave_i24_wav_inter_f32"].view(np.uint32))
    assert code() == 0 and code() == 4
    if ndev >= 2:
        # 10. the Convolver class on ndev GPUs
        N = 8
        xs = [lcg_noise(6000, 100 + i) for i in range(N)]
        irs = [[decaying(3000, 200 + o * N + i) for i in range(N)] for o in range(N)]
        assert [code() for _ in range(N * N)] == [0] * (N * N)
        assert code() == 2                                                # exchange fused into the inverse-FFT kernels
        for o in range(N):
            truth = sum(ck.direct_convolve_delayed(irs[o][i], xs[i], 256) for i in range(N))
            assert ck.rel_rms(take(6000), truth) <= TOL32
        N = 4
        xs = [lcg_noise(4096, 300 + i) for i in range(N)]
        irs = [[decaying(9000, 400 + o * N + i) for i in range(N)] for o in range(N)]
        assert [code() for _ in range(N * N)] == [0] * (N * N)
        assert code() == 1                                                # kLatencyShort: owner-side peer reads
        for o in range(N):
            truth = sum(ck.direct_convolve_delayed(irs[o][i], xs[i], 128) for i in range(N))
            assert ck.rel_rms(take(4096), truth) <= TOL32
        xs = [lcg_noise(4096, 500 + i) for i in range(N)]
        irs = [decaying(3000, 600 + i) for i in range(N)]
        assert [code() for _ in range(N)] == [0] * N
        assert code() == 0                                                # parallel channels: nothing crosses
        for i in range(N):
            assert ck.rel_rms(take(4096), ck.direct_convolve_delayed(irs[i], xs[i], 512)) <= TOL32
    assert pos[0] == len(data)
