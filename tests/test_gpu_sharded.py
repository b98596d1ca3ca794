"""Multi-GPU path on real devices (skipped on a box with fewer than 2 GPUs): two ranks under torch.distributed /
NCCL, input channels sharded, once with the NCCL reduce-scatter and once with the exchange fused into the
inverse-FFT epilogue (peer stores).  Both must equal the single-GPU result up to summation order (1e-6)."""
import os
import socket

import numpy as np
import pytest
import torch

import checkers as ck

pytestmark = pytest.mark.gpu

N_IN, N_OUT, L, B, BLOCKS = 8, 8, 3000, 256, 20


def _inputs():
    irs = [[ck.synth_ir(L, 1100 + 10 * o + i) for i in range(N_IN)] for o in range(N_OUT)]
    xs = np.stack([ck.synth_audio(B * BLOCKS, 1100 + i) for i in range(N_IN)])
    return irs, xs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir, exchange):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from hisstools_library_b200.sharded import ShardedConvolver
    irs, xs = _inputs()
    cv = ShardedConvolver(N_IN, N_OUT, False, 2 * B, maxLength=L, device=rank, exchange=exchange)
    assert cv.exchange == exchange
    cv.setResetOffset(0)
    for o in range(N_OUT):
        for i in range(N_IN):
            assert int(cv.set(i, o, irs[o][i], L, False)) == 0
    plan = cv.plan
    stream = torch.cuda.Stream()
    got = np.zeros((plan.local_outs, B * BLOCKS), np.float32)
    with torch.cuda.stream(stream):
        pos = 0
        for nb in [1, 1, 3, 1, 2, 4, 1, 7]:                  # hop-aligned calls of 1..7 hops
            n = nb * B
            x_local = torch.from_numpy(np.ascontiguousarray(xs[plan.in_lo:plan.in_hi, pos:pos + n])).cuda()
            y_shard = torch.zeros(plan.local_outs, n, device="cuda")
            assert cv.process_device(x_local, y_shard, n, stream.cuda_stream)
            stream.synchronize()
            got[:, pos:pos + n] = y_shard.cpu().numpy()
            pos += n
        assert pos == B * BLOCKS
    np.save(os.path.join(out_dir, "%s_rank%d.npy" % (exchange, rank)), got)
    dist.barrier()
    cv.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("exchange", ["nccl", "fused"])
def test_multi_gpu_sharded_matrix(tmp_path, exchange, world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), exchange), nprocs=world, join=True)
    irs, xs = _inputs()
    got = np.concatenate([np.load(tmp_path / ("%s_rank%d.npy" % (exchange, r))) for r in range(world)], axis=0)
    import hisstools_library_b200 as hb
    cv = hb.Convolver(N_IN, N_OUT, False, 2 * B, maxLength=L)
    cv.setResetOffset(0)
    for o in range(N_OUT):
        for i in range(N_IN):
            cv.set(i, o, irs[o][i], L, False)
    single = np.zeros((N_OUT, B * BLOCKS), np.float32)
    cv.process(xs, single, N_IN, N_OUT, B * BLOCKS)
    for o in range(N_OUT):
        assert ck.rel_rms(got[o], single[o]) <= 1e-6
        truth = sum(ck.direct_convolve_delayed(irs[o][i], xs[i], B) for i in range(N_IN))
        assert ck.rel_rms(got[o], truth) <= 1e-5


# ---- config 4 at its stated size through the sharded engine, against the compiled reference ---------------------------

C4_INS, C4_OUTS, C4_B, C4_P = 64, 64, 4096, 64


def _c4_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from hisstools_library_b200.sharded import ShardedConvolver
    taps, hops = C4_B * C4_P, C4_P + 16
    cv = ShardedConvolver(C4_INS, C4_OUTS, False, 2 * C4_B, maxLength=taps, device=rank, exchange="auto")
    assert cv.exchange == "fused"
    cv.setResetOffset(0)
    plan = cv.plan
    eng = cv.engine.m.tail
    gen = torch.Generator(device=dev)
    decay = torch.exp(-6.9 * torch.arange(taps, device=dev, dtype=torch.float64) / taps).float()
    for o in range(C4_OUTS):
        for i in range(plan.local_ins):
            ir = bench.device_ir(gen, bench.ir_seed(C4_INS, C4_OUTS, 0, o, plan.in_lo + i), taps, decay, torch.float32)
            assert eng.set_ir_device(0, i, o, ir.data_ptr(), taps) == 0
    pool = bench.input_pool(gen, rank, plan.local_ins, C4_B, 4, torch.float32, dev)
    stream = torch.cuda.Stream(device=dev)
    keep = [0, plan.local_outs - 1]
    xs = torch.cat([pool[k % 4] for k in range(hops)], dim=1).contiguous()
    # one block per call (the timed path), then calls of several blocks: multi-hop batches on the sharded engine (one pass of
    # this rank's IR spectra for up to 8 hops, one inverse launch and one owner-side sum for the whole batch)
    for tag, calls in (("", [1]), ("_b8", [8]), ("_mixed", [4, 1, 2, 8, 3])):
        cv.reset()
        got = torch.zeros(2, hops * C4_B, device=dev)
        with torch.cuda.stream(stream):
            pos, k = 0, 0
            while pos < hops * C4_B:
                m = min(calls[k % len(calls)] * C4_B, hops * C4_B - pos)
                yb = torch.zeros(plan.local_outs, m, device=dev)
                assert cv.process_device(xs[:, pos:pos + m].contiguous(), yb, m, stream.cuda_stream)
                for q, r in enumerate(keep):
                    got[q, pos:pos + m].copy_(yb[r], non_blocking=True)
                pos += m
                k += 1
        torch.cuda.synchronize()
        assert eng.shard_status() == 0
        np.save(os.path.join(out_dir, "c4%s_rank%d.npy" % (tag, rank)), got.cpu().numpy())
    dist.barrier()
    cv.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_config4_full_size_against_reference(tmp_path, world):
    """BASELINE config 4 (64 x 64, 262144 taps, 4096-sample blocks) with its inputs sharded over every rank count the box
    has, fused exchange, P + 16 blocks: the first and last output row of the first and of the last rank against the
    reference's rows of 64 MonoConvolves (bench.py's synthetic data: what its SCALE lines print as parity)."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    if ck.ref() is None:
        pytest.skip("compiled reference not shipped")
    import bench
    import torch.multiprocessing as mp
    mp.spawn(_c4_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    taps, hops = C4_B * C4_P, C4_P + 16
    own, l_ins = C4_OUTS // world, C4_INS // world
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    decay = torch.exp(-6.9 * torch.arange(taps, device=dev, dtype=torch.float64) / taps).float()
    rows = [0, own - 1, (world - 1) * own, C4_OUTS - 1]
    irs = np.stack([np.stack([bench.device_ir(gen, bench.ir_seed(C4_INS, C4_OUTS, 0, o, i), taps, decay, torch.float32).cpu().numpy()
                              for i in range(C4_INS)]) for o in rows])
    xs = []
    for r in range(world):
        pool = bench.input_pool(gen, r, l_ins, C4_B, 4, torch.float32, dev)
        xs.append(torch.cat([pool[k % 4] for k in range(hops)], dim=1).cpu().numpy())
    want = ck.ref_matrix_run(irs, np.concatenate(xs, axis=0), 2 * C4_B)
    for tag in ("", "_b8", "_mixed"):
        first, last = np.load(tmp_path / ("c4%s_rank0.npy" % tag)), np.load(tmp_path / ("c4%s_rank%d.npy" % (tag, world - 1)))
        for q, got in enumerate([first[0], first[1], last[0], last[1]]):
            assert ck.rel_rms(got, want[q]) <= 1e-5, (tag, q, ck.rel_rms(got, want[q]))
            assert ck.rel_rms(got[-16 * C4_B:], want[q][-16 * C4_B:]) <= 1e-5


# ---- a rank with nothing loaded must not stall or trap its peers -------------------------------------------------------

def _silent_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from hisstools_library_b200.sharded import ShardedConvolver
    irs, xs = _inputs()
    cv = ShardedConvolver(N_IN, N_OUT, False, 2 * B, maxLength=L, device=rank, exchange="fused")
    cv.setResetOffset(0)
    plan = cv.plan
    stream = torch.cuda.Stream()
    got = np.zeros((plan.local_outs, B * BLOCKS), np.float32)
    with torch.cuda.stream(stream):
        pos = 0
        for call, nb in enumerate([1, 2, 1, 3, 1, 4, 1, 7]):
            if call == 3:
                # from here on rank 0's inputs have their IRs; the LAST rank never loads any
                for o in range(N_OUT):
                    for i in range(N_IN):
                        if plan.input_owner(i) != world - 1:
                            cv.set(i, o, irs[o][i], L, False)
            n = nb * B
            x_local = torch.from_numpy(np.ascontiguousarray(xs[plan.in_lo:plan.in_hi, pos:pos + n])).cuda()
            y_shard = torch.full((plan.local_outs, n), 5.0, device="cuda")
            cv.process_device(x_local, y_shard, n, stream.cuda_stream)
            stream.synchronize()
            got[:, pos:pos + n] = y_shard.cpu().numpy()
            pos += n
    assert cv.engine.m.tail.shard_status() == 0
    np.save(os.path.join(out_dir, "silent_rank%d.npy" % rank), got)
    dist.barrier()
    cv.close()
    dist.destroy_process_group()


def test_fused_exchange_with_a_rank_that_has_no_ir(tmp_path):
    """One rank never loads an impulse response (and none has one during the first calls): it delivers silence instead of
    leaving its peers waiting, the hop sequence stays in step, nothing traps (hb_conv_shard_status stays 0)."""
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    mp.spawn(_silent_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    irs, xs = _inputs()
    got = np.concatenate([np.load(tmp_path / ("silent_rank%d.npy" % r)) for r in range(world)], axis=0)
    start = (1 + 2 + 1) * B                                   # the IRs arrive before the fourth call: the stream restarts there
    assert np.all(got[:, :start] == 0)
    for o in range(N_OUT):
        truth = sum(ck.direct_convolve_delayed(irs[o][i], xs[i][start:], B) for i in range(N_IN // 2))
        assert ck.rel_rms(got[o, start:], truth) <= 1e-5
