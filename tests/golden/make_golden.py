"""Generates tests/golden/golden.npz from the UNMODIFIED reference compiled in place
(oracle/_ref, via tests/checkers.py).  Run in a container that has /root/reference:

    make -C oracle ref && python tests/golden/make_golden.py

The fixtures pin the oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_*.py)
on machines where the reference itself is absent.  Inputs are stored with the outputs, so nothing
depends on a random generator being reproducible.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import checkers as ck  # noqa: E402


def main():
    rl, rs = ck.ref(), ck.ref_spectral()
    assert rl is not None and rs is not None, "build oracle/_ref first"
    g = {}
    rng = np.random.default_rng(20261017)

    # ---- FFT family -------------------------------------------------------------------------
    for dtype, suf in ((np.float32, "_f32"), (np.float64, "_f64")):
        setup = getattr(rl, "ref_fft_setup" + suf)(13)
        for log2n in (1, 2, 3, 4, 5, 6, 9, 12):
            n = 1 << log2n
            for op in ("fft", "ifft", "rfft", "rifft"):
                planes = n if op in ("fft", "ifft") else n >> 1
                if planes < 1:
                    continue
                re = rng.uniform(-1, 1, planes).astype(dtype)
                im = rng.uniform(-1, 1, planes).astype(dtype)
                key = "%s%s_%d" % (op, suf, log2n)
                g[key + "_in"] = np.stack([re, im])
                getattr(rl, "ref_%s%s" % (op, suf))(setup, ck.fptr(re), ck.fptr(im), log2n)
                g[key + "_out"] = np.stack([re, im])
        for log2n, in_length in ((5, 17), (8, 255), (10, 1024), (10, 700)):
            n = 1 << log2n
            x = rng.uniform(-1, 1, in_length).astype(dtype)
            re, im = np.zeros(n >> 1, dtype), np.zeros(n >> 1, dtype)
            getattr(rl, "ref_rfft_real" + suf)(setup, ck.fptr(x), ck.fptr(re), ck.fptr(im), in_length, log2n)
            key = "rfft_real%s_%d_%d" % (suf, log2n, in_length)
            g[key + "_in"] = x
            g[key + "_out"] = np.stack([re, im])
            back = np.zeros(n, dtype)
            getattr(rl, "ref_rifft_real" + suf)(setup, ck.fptr(re), ck.fptr(im), ck.fptr(back), log2n)
            g[key + "_back"] = back
        getattr(rl, "ref_fft_setup_free" + suf)(setup)

    # ---- PartitionedConvolve ------------------------------------------------------------------
    pconv_cases = [
        ("c1", 1024, 4096, 512, None, 0, 0, 0, 24),
        ("ragged", 256, 1000, 77, None, 0, 0, 0, 30),
        ("phase", 256, 1000, 129, None, 0, 0, 17, 30),
        ("slice", 512, 4000, 300, 4000, 1000, 1500, 0, 30),
        ("trunc", 512, 1281, 256, 1000, 0, 0, 0, 20),
        ("min", 32, 64, 16, None, 0, 0, 0, 40),
    ]
    for name, fft, ir_len, block, max_len, offset, length, reset_offset, hops in pconv_cases:
        ir = ck.synth_ir(ir_len, 11)
        x = ck.synth_audio(hops * (fft // 2) + 13, 12)
        y, err = ck.ref_pconv_run(fft, ir, x, block, max_len, offset, length, reset_offset)
        g["pconv_%s_ir" % name] = ir
        g["pconv_%s_x" % name] = x
        g["pconv_%s_y" % name] = y
        g["pconv_%s_meta" % name] = np.array([fft, block, -1 if max_len is None else max_len, offset, length,
                                              reset_offset, err], np.int64)

    # ---- MonoConvolve shipped latency modes (kLatencyZero/Short/Medium = 0/1/2) ---------------------
    ir = ck.synth_ir(9000, 21)
    x = ck.synth_audio(24576, 22)
    g["mono_ir"], g["mono_x"] = ir, x
    for mode in (0, 1, 2):
        h = rl.ref_mono_create_latency(len(ir), mode)
        rl.ref_mono_set_reset_offset(h, 0)
        assert rl.ref_mono_set(h, ck.fptr(ir), len(ir), 1) == 0
        rl.ref_mono_set_reset_offset(h, 0)
        y, tmp = np.zeros_like(x), np.zeros_like(x)
        pos = 0
        while pos < len(x):
            n = min(512, len(x) - pos)
            rl.ref_mono_process(h, ck.fptr(x[pos:]), ck.fptr(tmp), ck.fptr(y[pos:]), n, 0)
            pos += n
        rl.ref_mono_destroy(h)
        g["mono_y_mode%d" % mode] = y

    # ---- Convolver (3 in x 2 out, kLatencyShort, IRs longer than the default 16384 allocation) ----
    n_in, n_out, L, n = 3, 2, 17000, 256 * 80
    cv = rl.ref_conv_create(n_in, n_out, 1)
    irs = np.stack([ck.synth_ir(L, 30 + p) for p in range(n_in * n_out)]).reshape(n_out, n_in, L)
    for o in range(n_out):
        for i in range(n_in):
            assert rl.ref_conv_set_f32(cv, i, o, ck.fptr(irs[o, i]), L, 1) == 0
    xs = np.stack([ck.synth_audio(n, 40 + i) for i in range(n_in)])
    ys = np.zeros((n_out, n), np.float32)
    block = 256
    for pos in range(0, n, block):
        xb = np.ascontiguousarray(xs[:, pos:pos + block])
        yb = np.zeros((n_out, block), np.float32)
        rl.ref_conv_process_f32(cv, ck.planar_ptrs(xb), ck.planar_ptrs(yb), n_in, n_out, block)
        ys[:, pos:pos + block] = yb
    rl.ref_conv_destroy(cv)
    g["conv_irs"], g["conv_x"], g["conv_y"] = irs, xs, ys

    # ---- uniform N x M matrix (config-3 shape, scaled down): 4 -> 2, FFT 512 -------------------------
    n_in, n_out, L, fft, n = 4, 2, 3000, 512, 256 * 40
    m = rl.ref_matrix_create(n_in, n_out, L, fft, 0)
    irs = np.stack([ck.synth_ir(L, 50 + p) for p in range(n_in * n_out)]).reshape(n_out, n_in, L)
    for o in range(n_out):
        for i in range(n_in):
            rl.ref_matrix_set(m, i, o, ck.fptr(irs[o, i]), L)
    xs = np.stack([ck.synth_audio(n, 60 + i) for i in range(n_in)])
    ys = np.zeros((n_out, n), np.float32)
    rl.ref_matrix_process(m, ck.planar_ptrs(xs), ck.planar_ptrs(ys), n)
    rl.ref_matrix_destroy(m)
    g["matrix_irs"], g["matrix_x"], g["matrix_y"] = irs, xs, ys
    g["matrix_meta"] = np.array([fft], np.int64)

    # ---- double-precision partitioned convolution (restated loop on the reference's double FFT) --
    ird = ck.synth_ir(5000, 70).astype(np.float64)
    xd = ck.synth_audio(256 * 50, 71).astype(np.float64)
    h = rl.ref_restated_create_f64(512)
    rl.ref_restated_set_f64(h, ck.fptr(ird), len(ird))
    yd = np.zeros_like(xd)
    rl.ref_restated_process_f64(h, ck.fptr(xd), ck.fptr(yd), len(xd))
    rl.ref_restated_destroy_f64(h)
    g["pconv64_ir"], g["pconv64_x"], g["pconv64_y"] = ird, xd, yd

    # ---- spectral_processor::convolve, five edge modes -------------------------------------------
    for dtype, suf in ((np.float32, "_f32"), (np.float64, "_f64")):
        for n1, n2 in ((1000, 300), (300, 1000), (64, 64), (7, 2), (1, 9)):
            a = rng.uniform(-1, 1, n1).astype(dtype)
            b = rng.uniform(-1, 1, n2).astype(dtype)
            g["spec%s_%d_%d_a" % (suf, n1, n2)] = a
            g["spec%s_%d_%d_b" % (suf, n1, n2)] = b
            for mode in range(5):
                y = np.zeros(n1 + n2 + 8, dtype)
                size = getattr(rs, "ref_spectral_convolve" + suf)(ck.fptr(y), ck.fptr(a), n1, ck.fptr(b), n2, mode, 32768)
                g["spec%s_%d_%d_m%d" % (suf, n1, n2, mode)] = y[:size]

    # ---- spectral_processor::correlate (real) and the complex-input convolve / correlate --------------
    # complex Wrap / WrapCentre read past the transform in the reference: kept out of the fixtures
    for dtype, suf in ((np.float32, "_f32"), (np.float64, "_f64")):
        for n1, n2 in ((1000, 300), (300, 1000), (64, 64), (7, 2), (1, 9)):
            a, b = g["spec%s_%d_%d_a" % (suf, n1, n2)], g["spec%s_%d_%d_b" % (suf, n1, n2)]
            for mode in range(5):
                y = np.zeros(n1 + n2 + 8, dtype)
                size = getattr(rs, "ref_spectral_binary" + suf)(ck.fptr(y), ck.fptr(a), n1, ck.fptr(b), n2, mode, 1, 32768)
                g["corr%s_%d_%d_m%d" % (suf, n1, n2, mode)] = y[:size]
            ai = rng.uniform(-1, 1, max(n1 // 2, 1)).astype(dtype)          # shorter imaginary plane
            bi = rng.uniform(-1, 1, n2).astype(dtype)
            g["cspec%s_%d_%d_ai" % (suf, n1, n2)] = ai
            g["cspec%s_%d_%d_bi" % (suf, n1, n2)] = bi
            for op in (0, 1):
                for mode in (0, 3, 4):
                    yr, yi = np.zeros(n1 + n2 + 8, dtype), np.zeros(n1 + n2 + 8, dtype)
                    size = getattr(rs, "ref_spectral_binary_complex" + suf)(ck.fptr(yr), ck.fptr(yi), ck.fptr(a), n1, ck.fptr(ai), len(ai),
                                                                            ck.fptr(b), n2, ck.fptr(bi), n2, mode, op, 32768)
                    g["cspec%s_%d_%d_op%d_m%d" % (suf, n1, n2, op, mode)] = np.stack([yr[:size], yi[:size]])

    # ---- spectral_processor::change_phase: minimum / interpolated / linear / maximum phase ------------
    for dtype, suf in ((np.float32, "_f32"), (np.float64, "_f64")):
        for size in (2, 5, 300, 1024):
            x = (rng.standard_normal(size) * np.exp(-5.0 * np.arange(size) / size)).astype(dtype)
            g["phase%s_%d_x" % (suf, size)] = x
            for k, (phase, tm) in enumerate(((0.0, 1.0), (0.3, 1.0), (0.5, 1.0), (1.0, 1.0), (0.8, 2.0))):
                y = np.zeros(4 * size + 16, dtype)
                n = getattr(rs, "ref_spectral_change_phase" + suf)(ck.fptr(y), ck.fptr(x), size, phase, tm, 1 << 16)
                g["phase%s_%d_k%d" % (suf, size, k)] = y[:n]

    path = os.path.join(HERE, "golden.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
