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
