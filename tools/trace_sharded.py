"""Kernel timeline of the sharded (multi-GPU) hop on every rank: python tools/trace_sharded.py [world] [fused|nccl] [overlapped|serial]"""
import os, sys, socket
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

KIND = ["fwd", "head", "tail", "inv", "gath"]
INS, OUTS, TAPS, B = 64, 64, 262144, 4096


def worker(rank, world, port, exchange, sched):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from hisstools_library_b200.sharded import ShardedConvolver
    cv = ShardedConvolver(INS, OUTS, False, 2 * B, maxLength=TAPS, device=rank, exchange=exchange)
    eng = cv.engine.m.tail
    eng.set_reset_offset(0)
    eng.set_schedule(sched == "overlapped")
    ir = torch.randn(TAPS, device=dev)
    for o in range(OUTS):
        for i in range(INS // world):
            eng.set_ir_device(0, i, o, ir.data_ptr(), TAPS)
    x = torch.rand(INS // world, B, device=dev)
    y = torch.zeros(OUTS // world, B, device=dev)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    for _ in range(5):
        cv.process_device(x, y, B, st.cuda_stream)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    eng.set_trace(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(12):
        cv.process_device(x, y, B, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    tr, hop = eng.get_trace()
    dist.barrier()
    for r in range(world):
        if r == rank:
            print("rank %d: exchange %s, schedule %s, %.1f us per hop over 12 hops" % (rank, cv.exchange, eng.schedule, e0.elapsed_time(e1) * 1e3 / 12))
            t0 = None
            for h in range(hop - 6, hop - 1):
                for k in range(5):
                    ent, ext = tr[h % 16, k, 0], tr[h % 16, k, 1]
                    m = ent > 0
                    if not m.any():
                        continue
                    if t0 is None:
                        t0 = int(ent[m].min())
                    e, x_ = ent[m].astype(np.int64) - t0, ext[m].astype(np.int64) - t0
                    print("  hop %3d %-4s ctas %3d  entry %9.1f .. %9.1f us   exit %9.1f .. %9.1f us" % (h, KIND[k], int(m.sum()), e.min() / 1e3, e.max() / 1e3, x_.min() / 1e3, x_.max() / 1e3))
            sys.stdout.flush()
        dist.barrier()
    cv.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    exchange = sys.argv[2] if len(sys.argv) > 2 else "fused"
    sched = sys.argv[3] if len(sys.argv) > 3 else "overlapped"
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    import torch.multiprocessing as mp
    mp.spawn(worker, args=(world, port, exchange, sched), nprocs=world, join=True)
