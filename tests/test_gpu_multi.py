"""One Convolver object on several GPUs of ONE process (hb_matrix_create_multi, the `devices=` argument of the mirror and the
device-list constructors of include/HIRT_Multichannel_Convolution/Convolver.h): the reference's single process(ins, outs, ...)
call with host pointers, whatever the call sizes, against the same object on one GPU, float64 direct convolution and -- at
BASELINE config 4's full size -- the compiled reference.  Skipped on a box with fewer than 2 GPUs."""
import numpy as np
import pytest
import torch

import checkers as ck

pytestmark = pytest.mark.gpu

TOL32 = 1e-5


def _worlds():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return [w for w in (2, 4, 8) if w <= n]


def _stream(cv, xs, n_out, sizes):
    n_in, n = xs.shape
    ys = np.zeros((n_out, n), np.float32)
    pos, k = 0, 0
    while pos < n:
        m = min(sizes[k % len(sizes)], n - pos)
        yb = np.zeros((n_out, m), np.float32)
        cv.process(np.ascontiguousarray(xs[:, pos:pos + m]), yb, n_in, n_out, m)
        ys[:, pos:pos + m] = yb
        pos += m
        k += 1
    return ys


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("scheme,exchange,delay", [((False, 512), "fused", 256), (("kLatencyShort",), "peer-reads", 128), (("kLatencyZero",), "peer-reads", 0)])
def test_multi_device_matrix_matches_one_device(world, scheme, exchange, delay):
    if world not in _worlds():
        pytest.skip("needs %d GPUs" % world)
    import hisstools_library_b200 as hb
    n_in = n_out = 8
    L = 5000
    B = 256
    scheme = tuple(getattr(hb, s) if isinstance(s, str) else s for s in scheme)
    irs = [[ck.synth_ir(L, 1300 + 10 * o + i) for i in range(n_in)] for o in range(n_out)]
    xs = np.stack([ck.synth_audio(B * 30 + 77, 1300 + i) for i in range(n_in)])
    sizes = [B, B, 100, 3 * B, 1, 2 * B + 5, B, 4 * B]
    outs = {}
    for devices in (None, list(range(world))):
        cv = hb.Convolver(n_in, n_out, *scheme, maxLength=L, devices=devices)
        cv.setResetOffset(0)
        y = np.full((n_out, 64), 3.0, np.float32)
        cv.process(xs[:, :64].copy(), y, n_in, n_out, 64)
        assert np.all(y == 0)                                # nothing loaded: the Convolver class writes silence (Convolver.cpp:146-153)
        for o in range(n_out):
            for i in range(n_in):
                assert cv.set(i, o, irs[o][i], L, False) == 0
        assert cv.set(n_in, 0, irs[0][0], L, False) == hb.CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE
        if devices is not None:
            assert cv.matrix.exchange == exchange
        outs[devices is None] = _stream(cv, xs, n_out, sizes)
        # a reset restarts the stream from silence on every device
        cv.reset()
        again = _stream(cv, xs[:, :B * 6], n_out, [B])
        assert ck.rel_rms(again, outs[devices is None][:, :B * 6]) <= 1e-6
    for o in range(n_out):
        assert ck.rel_rms(outs[False][o], outs[True][o]) <= 2e-6
        truth = sum(ck.direct_convolve_delayed(irs[o][i], xs[i], delay) for i in range(n_in))
        assert ck.rel_rms(outs[False][o], truth) <= TOL32


def test_multi_device_matrix_with_partitions_above_one_cta():
    """A uniform scheme of 65536-point FFTs (32768-sample partitions: above the one-CTA transform limit, so the transforms are
    four-step chains) keeps the fused exchange: the chain leaves this rank's partial blocks in a local buffer and k_shard_deliver
    hands them to their owners.  4 x 4 matrix on two GPUs against one GPU and float64 direct convolution, ragged calls."""
    if 2 not in _worlds():
        pytest.skip("needs 2 GPUs")
    import hisstools_library_b200 as hb
    n_in = n_out = 4
    B, L = 32768, 70000
    irs = [[ck.synth_ir(L, 1500 + 10 * o + i) for i in range(n_in)] for o in range(n_out)]
    xs = np.stack([ck.synth_audio(B * 4 + 1234, 1500 + i) for i in range(n_in)])
    sizes = [B, 5000, B - 5000, 2 * B, 1234]
    outs = {}
    for devices in (None, [0, 1]):
        cv = hb.Convolver(n_in, n_out, False, 2 * B, maxLength=L, devices=devices)
        cv.setResetOffset(0)
        for o in range(n_out):
            for i in range(n_in):
                assert cv.set(i, o, irs[o][i], L, False) == 0
        if devices is not None:
            assert cv.matrix.exchange == "fused"
        outs[devices is None] = _stream(cv, xs, n_out, sizes)
    for o in range(n_out):
        assert ck.rel_rms(outs[False][o], outs[True][o]) <= 2e-6
        truth = sum(ck.direct_convolve_delayed_fft(irs[o][i], xs[i], B) for i in range(n_in))
        assert ck.rel_rms(outs[False][o], truth) <= TOL32


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_device_parallel_banks_and_double(world):
    """Convolver(numIO, ...) dealt to the GPUs bank by bank (no exchange), double engine included."""
    if world not in _worlds():
        pytest.skip("needs %d GPUs" % world)
    import hisstools_library_b200 as hb
    K, B, L = 8, 128, 1500
    irs = [ck.synth_ir(L, 1400 + k) for k in range(K)]
    xs = np.stack([ck.synth_audio(B * 20 + 9, 1400 + k) for k in range(K)])
    for dtype, tol in ((np.float32, TOL32), (np.float64, 1e-12)):
        cv = hb.Convolver(K, False, 2 * B, maxLength=L, devices=list(range(world)), dtype=dtype)
        cv.setResetOffset(0)
        for k in range(K):
            assert cv.set(k, k, irs[k].astype(dtype), L, False) == 0
        assert cv.matrix.exchange == "none"
        x = xs.astype(dtype)
        y = np.zeros_like(x)
        cv.process(x, y, K, K, x.shape[1])
        for k in range(K):
            truth = ck.direct_convolve_delayed(irs[k].astype(dtype), x[k], B)
            assert ck.rel_rms(y[k], truth) <= tol * (1 if dtype == np.float32 else 10)


def test_multi_device_config4_full_size_against_reference():
    """BASELINE config 4 at its stated size behind ONE object and ONE host-pointer call on every GPU of the box: P + 16
    blocks, one block per call (the e2e path of bench.py at N GPUs) and a ragged tail, four output rows against the
    reference's rows of 64 MonoConvolves."""
    worlds = _worlds()
    if not worlds:
        pytest.skip("needs at least 2 GPUs")
    if ck.ref() is None:
        pytest.skip("compiled reference not shipped")
    import bench
    import hisstools_library_b200 as hb
    world = worlds[-1]
    ins = outs = 64
    B, P = 4096, 64
    taps, hops = B * P, P + 16
    cv = hb.Convolver(ins, outs, False, 2 * B, maxLength=taps, devices=list(range(world)))
    cv.setResetOffset(0)
    assert cv.matrix.exchange == "fused"
    engines = cv.matrix.shard_engines()
    l_ins = ins // world
    rows = [0, outs // world - 1, outs // 2, outs - 1]
    kept = {}
    for d in range(world):
        dev = torch.device("cuda", d)
        gen = torch.Generator(device=dev)
        decay = torch.exp(-6.9 * torch.arange(taps, device=dev, dtype=torch.float64) / taps).float()
        for o in range(outs):
            for i in range(l_ins):
                gi = d * l_ins + i
                ir = bench.device_ir(gen, bench.ir_seed(ins, outs, 0, o, gi), taps, decay, torch.float32)
                assert engines[d].set_ir_device(0, i, o, ir.data_ptr(), taps) == 0
                if o in rows:
                    kept[(o, gi)] = ir.cpu().numpy()
    xs = np.stack([ck.synth_audio(B * hops, 1500 + i) for i in range(ins)])
    ys = np.zeros((outs, B * hops), np.float32)
    for k in range(hops - 6):
        cv.process(xs[:, k * B:(k + 1) * B].copy(), ys[:, k * B:(k + 1) * B], ins, outs, B)
    pos = (hops - 6) * B
    for m in (B + 100, 2 * B - 100, 3 * B):                     # ragged and multi-block calls
        yb = np.zeros((outs, m), np.float32)
        cv.process(np.ascontiguousarray(xs[:, pos:pos + m]), yb, ins, outs, m)
        ys[:, pos:pos + m] = yb
        pos += m
    assert pos == B * hops
    irs = np.stack([np.stack([kept[(o, i)] for i in range(ins)]) for o in rows])
    want = ck.ref_matrix_run(irs, xs, 2 * B)
    for q, o in enumerate(rows):
        assert ck.rel_rms(ys[o], want[q]) <= TOL32, (o, ck.rel_rms(ys[o], want[q]))
        assert ck.rel_rms(ys[o][-16 * B:], want[q][-16 * B:]) <= TOL32
