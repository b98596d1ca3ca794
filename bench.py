#!/usr/bin/env python
"""bench.py -- throughput of the partitioned-convolution hot path on B200, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c3|c2|c1|c5]

A *step* is one pass of the hot path over one block of synthetic input: `--hops` hops (default 1) of
B samples on every input channel -> forward FFTs, the frequency-domain multiply-accumulate against
all partitioned IR spectra, inverse FFTs, scale and store of B samples on every output channel.

Default workload (N = 1) is BASELINE.json's headline configuration, config 4: Convolver 64-in x 64-out,
262144-tap IRs, 4096-sample blocks, float.  It fits one B200 (8 GiB of IR spectra).  For N > 1 the
input channels are sharded over the ranks (one process per GPU, `torch.distributed` / NCCL): every
rank convolves its inputs against all outputs and the partial output blocks are summed with one
reduce-scatter per step -- the cross-device form of NToMonoConvolve.cpp:39-42.  Total
work is fixed as N grows ("scaling": "strong").

Printed JSON (one line, rank 0):
  value      whole-job M output-samples/s, inputs resident in HBM, CUDA events on the launching stream,
             max over ranks
  e2e        the same metric through the host-pointer C-ABI call (hb_conv_process; host<->device
             copies inside the timed region)
  roofline   the dominant multiply-accumulate launch (the tail launch of the overlapped schedule):
             its algorithmic bytes / its mean duration measured with CUDA events inside the library on the
             stream it is launched on, against MEASURED_PEAKS.json; hop_frac = bytes per hop / whole hop period
  multi_hop_reuse  the same engine fed 4 blocks per call (every IR spectrum read once per call): reported
             separately, the per-hop byte figure does not apply to it
  cpu_baseline  the unmodified reference (oracle/_ref, compiled from /root/reference by oracle/Makefile)
             timed on this box's host cores on a bounded sample of the same workload (rank 0, N = 1)

`--impl reference` times only the reference's CPU implementation (all host threads) and prints the
same line with "impl": "reference".  oracle/ is used here strictly as the measured baseline / checker;
the product path (hisstools_library_b200) never touches it.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "Msamples/sec partitioned convolution (per-GPU and 8-GPU) vs HBM roofline"
UNIT = "Msamples/s"

# name -> (ins, outs, groups, taps, hop B, dtype, description)
WORKLOADS = {
    "c1": (1, 1, 1, 4096, 512, "f32", "MonoConvolve 1ch, 4096-tap IR, 512-sample blocks, float"),
    "c2": (1, 1, 1, 65536, 1024, "f32", "PartitionedConvolve 1ch, 65536-tap IR, 1024-sample blocks, float"),
    "c3": (8, 1, 1, 131072, 2048, "f32", "NToMonoConvolve 8-in->1-out, 131072-tap IRs, 2048-sample blocks, float"),
    "c4": (64, 64, 1, 262144, 4096, "f32", "Convolver 64-in x 64-out, 262144-tap IRs, 4096-sample blocks, float"),
    "c5": (1, 1, 16, 1048576, 8192, "f64", "PartitionedConvolve double: 16ch, 1M-tap IR, 8192-sample blocks"),
    # one rank's share of config 4 when its inputs are sharded over 2 / 4 / 8 GPUs, runnable on ONE GPU (ncu cannot wrap a
    # multi-rank command): the source of profiles/traffic.json's c4_n2 / c4_n4 / c4_n8 entries
    "c4r2": (32, 64, 1, 262144, 4096, "f32", "one rank of config 4 sharded over 2 GPUs: 32-in x 64-out, 262144-tap IRs, 4096-sample blocks"),
    "c4r4": (16, 64, 1, 262144, 4096, "f32", "one rank of config 4 sharded over 4 GPUs: 16-in x 64-out, 262144-tap IRs, 4096-sample blocks"),
    "c4r8": (8, 64, 1, 262144, 4096, "f32", "one rank of config 4 sharded over 8 GPUs: 8-in x 64-out, 262144-tap IRs, 4096-sample blocks"),
}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def committed_traffic(workload, n_gpus, overlapped):
    """DRAM bytes per launch of the multiply-accumulate kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("%s_n%d%s" % (workload, n_gpus, "_tail" if overlapped else ""))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        res = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return res
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                except ValueError:
                    continue
                for name, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            res.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        res["reasons"] = sorted(reasons)
        return res


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------------------
def reference_rate(workload, steps, warmup, hops_per_step, budget_s=None):
    """Times the compiled reference on a bounded sample of `workload`.
    Returns dict(value=M output-samples/s, cores, kind, sample, ms_per_step)."""
    import checkers as ck
    ins, outs, groups, taps, B, dtype, _ = WORKLOADS[workload]
    lib = ck.ref()
    if lib is None:
        raise RuntimeError("oracle/_ref/libhisstools_ref.so is missing (built by oracle/Makefile where /root/reference exists)")
    threads = int(lib.ref_hardware_threads()) or 1
    n = B * hops_per_step
    rng = np.random.default_rng(7)
    if dtype == "f64":
        # no double convolver class exists in the reference: the restated double loop over the
        # reference's own double FFT (oracle/ref_shim.cpp PConvRestated<double>), one object per channel
        chans = min(groups, max(threads, 1))
        objs = (C.c_void_p * chans)()
        decay = np.exp(-6.9 * np.arange(taps) / taps)
        for c in range(chans):
            objs[c] = lib.ref_restated_create_f64(2 * B)
            ir = rng.standard_normal(taps) * decay
            lib.ref_restated_set_f64(objs[c], ck.fptr(ir), taps)
        x = rng.uniform(-1, 1, (chans, n))
        y = np.zeros((chans, n))
        xp, yp = ck.planar_ptrs(x), ck.planar_ptrs(y)
        use = min(threads, chans)
        run = lambda w, h: lib.ref_restated_time_f64(objs, chans, xp, yp, n, w, h, use)
        rows, sample = chans, "%d of %d channels (independent objects), restated double loop on the reference FFT" % (chans, groups)
        cleanup = lambda: [lib.ref_restated_destroy_f64(objs[c]) for c in range(chans)]
        kind = "reference"
    else:
        rows = outs if outs <= 8 else int(min(outs, max(8, min(threads, 32))))
        m = lib.ref_matrix_create(ins, rows, taps, 2 * B, 0)
        if not m:
            raise RuntimeError("reference matrix allocation failed")
        decay = np.exp(-6.9 * np.arange(taps) / taps)
        for o in range(rows):
            for i in range(ins):
                ir = (rng.standard_normal(taps) * decay).astype(np.float32)
                lib.ref_matrix_set(m, i, o, ck.fptr(ir), taps)
        x = rng.uniform(-1, 1, (ins, n)).astype(np.float32)
        y = np.zeros((rows, n), np.float32)
        xp, yp = ck.planar_ptrs(x), ck.planar_ptrs(y)
        use = min(threads, rows)
        run = lambda w, h: lib.ref_matrix_time(m, xp, yp, n, w, h, use)
        sample = "%d of %d output rows x all %d inputs (rows are independent objects in the reference)" % (rows, outs, ins) if rows < outs \
            else "full workload"
        cleanup = lambda: lib.ref_matrix_destroy(m)
        kind = "reference"
    # the reference only multiplies the partitions whose FDL slot has been filled (mValidPartitions,
    # PartitionedConvolve.cpp:285,373): warm up until the delay line is full, or early steps are cheap
    P = (taps + B - 1) // B
    warmup = max(warmup, (P + hops_per_step) // hops_per_step + 1)
    if budget_s is not None:
        t1 = run(warmup, 2) / 2
        steps = int(max(3, min(400, budget_s / max(t1, 1e-6))))
        warmup = 0
    secs = run(warmup, steps)
    cleanup()
    return {"value": rows * n * steps / secs / 1e6, "unit": UNIT, "cores": use, "kind": kind,
            "sample": "%s; %d steps of %d samples, %d host threads, SSE2 -O2 build" % (sample, steps, n, use),
            "ms_per_step": secs / steps * 1e3, "steps": steps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        r = reference_rate(args.workload, args.steps, args.warmup, args.hops)
    except Exception as e:                                          # the oracle always exists; report why it did not run
        print(json.dumps({"impl": "reference", "unavailable": str(e)}))
        return
    ins, outs, groups, taps, B, dtype, desc = WORKLOADS[args.workload]
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "impl": "reference",
            "config": {"workload": desc, "hops_per_step": args.hops, "sample": r["sample"]},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import hisstools_library_b200 as hb
    from hisstools_library_b200 import _abi
    from hisstools_library_b200.convolve import _Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        sys.stderr.write("bench.py: --gpus %d but WORLD_SIZE %d; using WORLD_SIZE\n" % (args.gpus, world))

    ins, outs, groups, taps, B, dtype, desc = WORKLOADS[args.workload]
    tdt = torch.float64 if dtype == "f64" else torch.float32
    ndt = np.float64 if dtype == "f64" else np.float32
    es = 8 if dtype == "f64" else 4
    # sharding: input channels of the matrix (collective = sum of partial outputs); independent banks otherwise
    if world > 1 and ins % world == 0 and ins >= world:
        mode, l_ins, l_groups = "inputs", ins // world, groups
    elif world > 1 and groups % world == 0 and groups >= world:
        mode, l_ins, l_groups = "groups", ins, groups // world
    elif world > 1:
        mode, l_ins, l_groups = "replicas", ins, groups
    else:
        mode, l_ins, l_groups = "single", ins, groups
    n = B * args.hops
    P = (taps + B - 1) // B

    sharded = None
    if mode == "inputs":
        # the public multi-GPU class: inputs sharded over the ranks, reduce-scatter of partial outputs
        from hisstools_library_b200.sharded import ShardedConvolver
        sharded = ShardedConvolver(ins, outs, False, 2 * B, maxLength=taps, dtype=ndt, device=local, exchange=args.exchange)
        eng = sharded.engine.m.tail
    else:
        eng = _Engine(ndt, l_groups, l_ins, outs, 2 * B, taps, 0, 0, local)
    eng.set_reset_offset(0)
    if args.variant is not None:
        eng.set_tuning(args.ctas_per_sm, args.variant)
    if args.schedule is not None:
        eng.set_schedule(args.schedule == "overlapped")
    if args.fft_path:
        eng.set_fft_path(args.fft_path)
    gen = torch.Generator(device=dev)
    decay = torch.exp(-6.9 * torch.arange(taps, device=dev, dtype=torch.float64) / taps).to(tdt)
    for g in range(l_groups):
        for o in range(outs):
            for i in range(l_ins):
                gi = (rank * l_ins + i) if mode == "inputs" else i
                gg = (rank * l_groups + g) if mode == "groups" else g
                gen.manual_seed(2000 + (gg * outs + o) * ins + gi)
                ir = torch.randn(taps, generator=gen, device=dev, dtype=tdt) * decay
                eng.set_ir_device(g, i, o, ir.data_ptr(), taps)
    torch.cuda.synchronize()
    assert eng.partitions == P, (eng.partitions, P)

    rows_in, rows_out = l_groups * l_ins, l_groups * outs
    n_pool = 4
    gen.manual_seed(1000 + rank)
    x_pool = [torch.rand(rows_in, n, generator=gen, device=dev, dtype=tdt) * 2 - 1 for _ in range(n_pool)]
    y_part = torch.zeros(rows_out, n, device=dev, dtype=tdt)
    shard_rows = rows_out // world if mode == "inputs" else rows_out
    y_shard = torch.zeros(shard_rows, n, device=dev, dtype=tdt) if mode == "inputs" else y_part
    # a real (non-default) stream: handle 0 would mean "the engine's own stream" to the C ABI
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def step(k):
        x = x_pool[k % n_pool]
        if sharded is not None:
            sharded.process_device(x, y_shard, n, stream.cuda_stream)
        else:
            eng.process_device(x.data_ptr(), n, y_part.data_ptr(), n, n, False, stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(max(args.warmup, 3)):
        step(k)
    barrier()

    # ---- device-resident timed region -------------------------------------------------------------
    lib = _abi.lib()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.hb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for k in range(args.steps):
        step(k)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.hb_launch_count() - launches0
    # ---- second pass over the same steps with CUDA events around every kernel launch (roofline figure);
    # kept out of the timed region above because the event records open small gaps between the kernels
    eng.set_profiling(True)
    for k in range(args.steps):
        step(k)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    prof, hops = eng.get_profile()
    eng.set_profiling(False)
    overlapped = eng.schedule == "overlapped"
    fused = eng.schedule == "fused"
    # the dominant launch: the tail multiply-accumulate (partitions 1..P-1, second stream) in the overlapped schedule,
    # the one multiply-accumulate over all partitions in the serial schedule
    ms_dom = prof["tail"] if overlapped else (prof["forward"] if fused else prof["cmac"])
    ms_head = prof["cmac"] if overlapped else 0.0
    # the inverse-FFT launch sits behind the wait for the tail; the event between the two is not ordered after the
    # wait, so only their sum is meaningful
    t = torch.tensor([ms, ms_dom / max(hops, 1), prof["forward"] / max(hops, 1), (prof["wait"] + prof["inverse"]) / max(hops, 1),
                      ms_head / max(hops, 1)], device=dev, dtype=torch.float64)
    cnt = torch.tensor([float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms, cmac_ms, fwd_ms, inv_ms, head_ms = [float(v) for v in t.tolist()]
    launches = int(cnt.item())

    # whole-job output samples per step: replicas each produce the full workload
    job_rows = groups * outs * (world if mode == "replicas" else 1)
    value = job_rows * n * args.steps / (ms * 1e-3) / 1e6

    # ---- multi-hop reuse: blocks of 4 hops per call, every IR spectrum streamed once per call (reported separately: the
    # per-hop byte figure above does not apply to it) --------------------------------------
    multi = None
    if sharded is None and args.hops == 1 and not args.no_multi_hop:
        mh = 4
        xm = [torch.rand(rows_in, n * mh, generator=gen, device=dev, dtype=tdt) * 2 - 1 for _ in range(2)]
        ym = torch.zeros(rows_out, n * mh, device=dev, dtype=tdt)
        for k in range(3):
            eng.process_device(xm[k % 2].data_ptr(), n * mh, ym.data_ptr(), n * mh, n * mh, False, stream.cuda_stream)
        barrier()
        m_steps = max(3, args.steps // 3)
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record(stream)
        for k in range(m_steps):
            eng.process_device(xm[k % 2].data_ptr(), n * mh, ym.data_ptr(), n * mh, n * mh, False, stream.cuda_stream)
        m1.record(stream)
        barrier()
        mms = m0.elapsed_time(m1) / m_steps
        multi = {"hops_per_call": mh, "value": job_rows * n * mh / (mms * 1e-3) / 1e6, "unit": UNIT, "ms_per_call": mms, "ms_per_hop": mms / mh,
                 "note": "hop-aligned calls of 4 blocks: on HBM-bound engines every IR spectrum is read once per call (k_cmac_tma_mh)"}
        del xm, ym

    # ---- end to end through the host-pointer boundary ----------------------------------------------
    e2e_steps = max(3, args.steps)
    e2e_warm = max(3, args.warmup)
    h2d = rows_in * n * es
    if world == 1:
        xin = [np.ascontiguousarray(x_pool[k].cpu().numpy()) for k in range(n_pool)]
        yout = np.zeros((rows_out, n), ndt)
        rows_y = [yout[r] for r in range(rows_out)]
        rows_x = [[xi[r] for r in range(rows_in)] for xi in xin]
        for k in range(e2e_warm):                                                         # staging buffers, copy streams, events
            eng.process(rows_x[k % n_pool], rows_y, n)
        torch.cuda.synchronize()
        per_call = []
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            # hb_conv_process: gathers the host rows, H2D, the hop's kernels, D2H, scatters the block to the host rows
            tc = time.perf_counter()
            eng.process(rows_x[k % n_pool], rows_y, n)
            per_call.append(time.perf_counter() - tc)
        torch.cuda.synchronize()                                                          # the last call's device work is inside the timed region
        e2e_s = time.perf_counter() - t0
        sys.stderr.write("e2e per-call ms: %s\n" % " ".join("%.3f" % (t * 1e3) for t in per_call))
        d2h = rows_out * n * es
    else:
        xh = [x_pool[k].cpu().pin_memory() for k in range(n_pool)]
        yh = torch.zeros(y_shard.shape, dtype=tdt).pin_memory()
        xd = torch.empty_like(x_pool[0])

        got = torch.cuda.Event()

        def e2e_step(k):
            xd.copy_(xh[k % n_pool], non_blocking=True)
            if sharded is not None:
                sharded.process_device(xd, y_shard, n, stream.cuda_stream)
            else:
                eng.process_device(xd.data_ptr(), n, y_part.data_ptr(), n, n, False, stream.cuda_stream)
            yh.copy_(y_shard, non_blocking=True)
            # the step is over when its result is in host memory; the tail of the NEXT hop, launched ahead on the engine's
            # second stream, keeps running behind this wait as it does in the device-resident loop (the final barrier +
            # synchronize below puts what is left of it inside the timed region)
            got.record(stream)
            got.synchronize()
        for k in range(e2e_warm):
            e2e_step(k)
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            e2e_step(k)
        torch.cuda.synchronize()
        barrier()
        e2e_s = time.perf_counter() - t0
        d2h = y_shard.numel() * es
        tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = job_rows * n * e2e_steps / e2e_s / 1e6

    # ---- roofline of the dominant kernel (multiply-accumulate), per rank ---------------------------
    peak, peak_src = measured_peak()
    bytes_per_hop = eng.bytes_per_hop                             # IR spectra + FDL + time-domain I/O of one hop
    bytes_per_launch = eng.bytes_per_launch                       # the dominant launch's share of it
    achieved = bytes_per_launch / (cmac_ms * 1e-3) / 1e9 if cmac_ms > 0 else 0.0
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": committed_traffic(args.workload, world, overlapped),
            "kernel": "k_cmac, tail launch: partitions 1..P-1 (frequency-domain multiply-accumulate)" if overlapped
                      else ("k_hop_fused (whole hop in one cluster launch)" if fused else "k_cmac (frequency-domain multiply-accumulate, all partitions)"),
            "bytes_per_launch": bytes_per_launch, "bytes_per_hop": bytes_per_hop, "kernel_ms": cmac_ms, "forward_fft_ms": fwd_ms,
            "head_cmac_ms": head_ms, ("wait_for_tail_plus_inverse_fft_ms" if overlapped else "inverse_fft_ms"): inv_ms,
            "kernel_share_of_step": cmac_ms * args.hops / (ms / args.steps), "peak_source": peak_src,
            "hop_frac": bytes_per_hop / ((ms / args.steps / args.hops) * 1e-3) / 1e9 / peak}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            r = reference_rate(args.workload, 0, 0, args.hops, budget_s=args.cpu_seconds)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if mode == "replicas" else "strong",
                "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": {"workload": desc, "hops_per_step": args.hops, "samples_per_step_per_channel": n, "partitions": P,
                           "sharding": mode, "local_inputs": l_ins, "outputs": outs, "groups": l_groups, "schedule": eng.schedule,
                           "transforms": {1: "one CTA each", 2: "cluster of 8 CTAs each (DSMEM)", 3: "four-step chains"}.get(eng.fft_path, "?"),
                           "l2": "inputs larger than L2: %.2f GiB of IR spectra per rank streamed every step" % (bytes_per_hop / 2 ** 30)
                                 if bytes_per_hop > 256e6 else "working set %.1f MiB is L2-resident (not an HBM-roofline case)" % (bytes_per_hop / 2 ** 20),
                           "collective": ("peer stores fused into the inverse-FFT epilogue (NVLink), owner-side sum" if sharded is not None and sharded.exchange == "fused"
                                          else "nccl reduce_scatter of partial output blocks") if mode == "inputs" else "none"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                        "api": "hb_conv_process (host pointers)" if world == 1 else
                               "pinned H2D + ShardedConvolver.process_device (%s) + D2H of this rank's output rows, waiting on the result event" %
                               ("hb_conv_process_shard_dev: peer stores from the inverse-FFT epilogue" if sharded is not None and sharded.exchange == "fused"
                                else "hb_matrix_process_dev + NCCL reduce_scatter")},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if multi is not None:
            line["multi_hop_reuse"] = multi
        print(json.dumps(line))
    if sharded is not None:
        sharded.close()
    else:
        eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--hops", type=int, default=1, help="hops (blocks of B samples) per step")
    ap.add_argument("--variant", type=int, default=None, help="multiply-accumulate kernel: 1 = TMA ring, 0 = direct loads")
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--schedule", default=None, choices=["overlapped", "serial"], help="hop schedule (default: the library's automatic choice)")
    ap.add_argument("--fft-path", type=int, default=0, choices=[0, 1, 2, 3],
                    help="transforms: 0 automatic, 1 one CTA each, 2 cluster of 8 CTAs each, 3 four-step chains (hb_conv_set_fft_path)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "fused", "nccl"], help="multi-GPU sum of partial outputs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-multi-hop", action="store_true", help="skip the multi-hop reuse leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
