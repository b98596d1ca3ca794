"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo process group, the sharding plan of
hisstools_library_b200.sharded and its exchange step (sum of partial output blocks), with the plain-C
oracle standing in for the CUDA engine on each rank (test infrastructure; the product default is CUDA).
The sharded result must equal the unsharded oracle result: only the summation order changes.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import checkers as ck


class OracleEngine:
    """stand-in with the engine interface of sharded._CudaMatrixEngine, backed by oracle/hiss_oracle.c"""

    def __init__(self, ins, outs, max_length, scheme, dtype, device):
        zero, A, B, C_, D = scheme
        self.lib = ck.oracle()
        self.ins, self.outs = ins, outs
        self.objs = [[self.lib.orc_mono_create_f32(max_length, int(zero), A, B, C_, D) for _ in range(ins)] for _ in range(outs)]
        self.loaded = False

    def set(self, i, o, ir, length, resize):
        ir = np.ascontiguousarray(ir, np.float32)
        self.loaded = True
        return self.lib.orc_mono_set_f32(self.objs[o][i], ck.fptr(ir), int(length), int(resize))

    def set_reset_offset(self, offset):
        pass

    def reset(self):
        for row in self.objs:
            for h in row:
                self.lib.orc_mono_reset_f32(h)

    def process_tensor(self, x, y, n, stream):
        if not self.loaded:
            return False
        xs = x.numpy()
        for o in range(self.outs):
            acc = np.zeros(n, np.float32)
            for i in range(self.ins):
                xi = np.ascontiguousarray(xs[i, :n])
                self.lib.orc_mono_process_f32(self.objs[o][i], ck.fptr(xi), ck.fptr(acc), n, 1)
            y[o, :n] = torch.from_numpy(acc)
        return True

    def close(self):
        for row in self.objs:
            for h in row:
                self.lib.orc_mono_destroy_f32(h)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


N_IN, N_OUT, L, B, BLOCKS = 4, 2, 700, 64, 12


def _inputs():
    irs = [[ck.synth_ir(L, 900 + 10 * o + i) for i in range(N_IN)] for o in range(N_OUT)]
    xs = np.stack([ck.synth_audio(B * BLOCKS, 900 + i) for i in range(N_IN)])
    return irs, xs


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hisstools_library_b200.sharded import ShardedConvolver
    import hisstools_library_b200 as hb
    irs, xs = _inputs()
    cv = ShardedConvolver(N_IN, N_OUT, False, 2 * B, maxLength=L, engine_factory=OracleEngine)
    plan = cv.plan
    assert (plan.in_lo, plan.in_hi) == (rank * N_IN // world, (rank + 1) * N_IN // world)
    codes = []
    for o in range(N_OUT):
        for i in range(N_IN):
            codes.append(int(cv.set(i, o, irs[o][i], L, False)))
    assert all(c == 0 for c in codes)
    assert cv.set(N_IN, 0, irs[0][0], L, False) == hb.CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE
    assert cv.set(0, N_OUT, irs[0][0], L, False) == hb.CONVOLVE_ERR_OUT_CHAN_OUT_OF_RANGE
    got = np.zeros((plan.local_outs, B * BLOCKS), np.float32)
    for b in range(BLOCKS):
        x_local = torch.from_numpy(np.ascontiguousarray(xs[plan.in_lo:plan.in_hi, b * B:(b + 1) * B]))
        y_shard = torch.zeros(plan.local_outs, B)
        assert cv.process_device(x_local, y_shard, B)
        got[:, b * B:(b + 1) * B] = y_shard.numpy()
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), got)
    cv.close()
    dist.barrier()
    dist.destroy_process_group()


def test_shard_plan():
    from hisstools_library_b200.sharded import ShardPlan
    p = ShardPlan(64, 64, 8, 3)
    assert (p.in_lo, p.in_hi, p.out_lo, p.out_hi, p.local_ins, p.local_outs) == (24, 32, 24, 32, 8, 8)
    assert p.input_owner(23) == 2 and p.input_owner(24) == 3 and p.output_owner(63) == 7
    with pytest.raises(ValueError):
        ShardPlan(6, 4, 4, 0)


def test_two_rank_sharded_matrix_equals_unsharded(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    irs, xs = _inputs()
    got = np.concatenate([np.load(tmp_path / ("rank%d.npy" % r)) for r in range(world)], axis=0)
    assert got.shape == (N_OUT, B * BLOCKS)
    for o in range(N_OUT):
        want = np.zeros(B * BLOCKS)
        for i in range(N_IN):
            y, _ = ck.oracle_pconv_run(2 * B, irs[o][i], xs[i], B)
            want += y
        assert ck.rel_rms(got[o], want) <= 1e-6
        truth = sum(ck.direct_convolve_delayed(irs[o][i], xs[i], B) for i in range(N_IN))
        assert ck.rel_rms(got[o], truth) <= 1e-5
