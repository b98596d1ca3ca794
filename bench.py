#!/usr/bin/env python
"""bench.py -- throughput of the partitioned-convolution hot path on B200, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c3|c2|c1|c5]

A *step* is one pass of the hot path over one batch of synthetic input: `blocks_per_step` consecutive blocks of
B samples on every input channel, fed ONE BLOCK PER CALL as the workload's block size says (forward FFTs, the
frequency-domain multiply-accumulate against all partitioned IR spectra, inverse FFTs, scale and store of B samples
on every output channel, per block).  blocks_per_step is chosen so that the K timed steps last at least ~0.6 s
(`--blocks-per-step` overrides it); every block is a separate call of the public entry point.

Default workload (N = 1) is BASELINE.json's headline configuration, config 4: Convolver 64-in x 64-out,
262144-tap IRs, 4096-sample blocks, float.  It fits one B200 (8 GiB of IR spectra).  For N > 1 the
input channels are sharded over the ranks (one process per GPU, `torch.distributed` / NCCL): every
rank convolves its inputs against all outputs and the partial output blocks are summed across the ranks --
the cross-device form of NToMonoConvolve.cpp:39-42.  Total work is fixed as N grows ("scaling": "strong").

Printed JSON (one line, rank 0):
  value      whole-job M output-samples/s, inputs resident in HBM, CUDA events on the launching stream with the
             engine's look-ahead stream joined before the closing event (steady state), max over ranks
  e2e        the same metric through the host-pointer C-ABI call (host<->device copies inside the timed region): hb_conv_process
             at N = 1; at N > 1 hb_matrix_process on ONE multi-device matrix (hb_matrix_create_multi) driven by rank 0 alone --
             the reference's one object / one process(ins, outs) call over all N GPUs -- while the other ranks wait on the host
             (the per-rank figure, pinned H2D + ShardedConvolver + D2H in every process, is kept as e2e.per_rank_processes)
  roofline   the dominant multiply-accumulate launch (the tail launch of the overlapped schedule):
             its algorithmic bytes / its mean duration measured with CUDA events inside the library on the
             stream it is launched on, against MEASURED_PEAKS.json; hop_frac = bytes per hop / whole hop period
  parity     relative RMS error of the benchmarked engine against the compiled reference (oracle/_ref) on the same
             synthetic stream, P + 16 blocks, computed OUTSIDE the timed regions (tolerance 1e-5 float, 1e-12 double)
  multi_hop_reuse  the same engine fed 4 / 8 blocks per call (every IR spectrum read once per call; launch-latency-bound
             engines: 8 / 64 blocks per call sharing their launches): reported separately, the per-hop byte figure does not apply
  cpu_baseline  the unmodified reference (oracle/_ref, compiled from /root/reference by oracle/Makefile)
             timed on this box's host cores on a bounded sample of the same workload (rank 0, N = 1)

`--impl reference` times only the reference's CPU implementation (all host threads) and prints the
same line with "impl": "reference".  oracle/ is used here strictly as the measured baseline / checker;
the product path (hisstools_library_b200) never touches it.
"""
import argparse
import ctypes as C
import json
import math
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
TOL = {"f32": 1e-5, "f64": 1e-12}


def workload_config(name):
    """the part of the JSON line that names the workload -- identical in both arms"""
    ins, outs, groups, taps, B, dtype, desc = WORKLOADS[name]
    return {"workload": desc, "inputs": ins, "outputs": outs, "independent_banks": groups, "taps": taps, "block_samples": B,
            "partitions": (taps + B - 1) // B, "element": dtype}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def committed_traffic(workload, n_gpus, overlapped):
    """DRAM bytes per launch of the multiply-accumulate kernel from the committed ncu capture (profiles/traffic.json:
    value, capture file and date), or None.  ncu cannot run inside a timed bench, so this is the one figure of the line
    that is read from a file; `traffic_source` names the capture it came from."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)
        key = "%s_n%d%s" % (workload, n_gpus, "_tail" if overlapped else "")
        src = d.get("_source", {}).get(key) or d.get("_source", {}).get("default")
        return d.get(key), src
    except Exception:
        return None, None


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


def mem_available_gib():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return float(line.split()[1]) / 2 ** 20
    except Exception:
        pass
    return 0.0


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------------------
def reference_rate(workload, steps, warmup, budget_s=None, full=False):
    """Times the compiled reference on `workload`, one block per step.  full = every output row when the host has the
    memory for it (the reference keeps 4 x taps floats per pair, PartitionedConvolve.cpp:84), else a bounded sample of rows
    (rows are independent objects in the reference).  Returns dict(value=M output-samples/s, cores, kind, sample, ms_per_step)."""
    import checkers as ck
    ins, outs, groups, taps, B, dtype, _ = WORKLOADS[workload]
    lib = ck.ref()
    if lib is None:
        raise RuntimeError("oracle/_ref/libhisstools_ref.so is missing (built by oracle/Makefile where /root/reference exists)")
    threads = int(lib.ref_hardware_threads()) or 1
    n = B
    rng = np.random.default_rng(7)
    decay = np.exp(-6.9 * np.arange(taps) / taps)
    if dtype == "f64":
        # no double convolver class exists in the reference: the restated double loop over the
        # reference's own double FFT (oracle/ref_shim.cpp PConvRestated<double>), one object per channel
        chans = groups if full else min(groups, max(threads, 1))
        objs = (C.c_void_p * chans)()
        for c in range(chans):
            objs[c] = lib.ref_restated_create_f64(2 * B)
            ir = rng.standard_normal(taps) * decay
            lib.ref_restated_set_f64(objs[c], ck.fptr(ir), taps)
        x = rng.uniform(-1, 1, (chans, n))
        y = np.zeros((chans, n))
        xp, yp = ck.planar_ptrs(x), ck.planar_ptrs(y)
        use = min(threads, chans)
        run = lambda w, h: lib.ref_restated_time_f64(objs, chans, xp, yp, n, w, h, use)
        rows = chans
        sample = ("full workload: all %d channels" % groups if chans == groups else "%d of %d channels" % (chans, groups)) + \
            " (independent objects), restated double loop on the reference FFT"
        cleanup = lambda: [lib.ref_restated_destroy_f64(objs[c]) for c in range(chans)]
    else:
        per_pair_gib = 4.0 * taps * 4 / 2 ** 30 * 1.15
        rows_fit = int((mem_available_gib() * 0.6) / max(per_pair_gib * ins, 1e-9))
        if full and rows_fit >= outs:
            rows = outs
        else:
            rows = outs if outs <= 8 else int(min(outs, max(8, min(threads, 32)), max(rows_fit, 1)))
        m = lib.ref_matrix_create(ins, rows, taps, 2 * B, 0)
        if not m:
            raise RuntimeError("reference matrix allocation failed")
        # a pool of distinct random IRs dealt to the pairs (throughput does not depend on the values)
        npool = min(64, ins * rows)
        pool = np.stack([(rng.standard_normal(taps) * decay).astype(np.float32) for _ in range(npool)])
        lib.ref_matrix_set_pool_mt(m, ck.planar_ptrs(pool), npool, taps, threads)
        x = rng.uniform(-1, 1, (ins, n)).astype(np.float32)
        y = np.zeros((rows, n), np.float32)
        xp, yp = ck.planar_ptrs(x), ck.planar_ptrs(y)
        use = min(threads, rows)
        run = lambda w, h: lib.ref_matrix_time(m, xp, yp, n, w, h, use)
        sample = "%d of %d output rows x all %d inputs (rows are independent objects in the reference)" % (rows, outs, ins) if rows < outs \
            else "full workload: all %d output rows x %d inputs" % (outs, ins)
        cleanup = lambda: lib.ref_matrix_destroy(m)
    # the reference only multiplies the partitions whose FDL slot has been filled (mValidPartitions,
    # PartitionedConvolve.cpp:285,373): warm up until the delay line is full, or early steps are cheap
    P = (taps + B - 1) // B
    warmup = max(warmup, P + 2)
    if budget_s is not None:
        t1 = run(warmup, 2) / 2
        steps = int(max(3, min(400, budget_s / max(t1, 1e-6))))
        warmup = 0
    secs = run(warmup, steps)
    cleanup()
    return {"value": rows * n * steps / secs / 1e6, "unit": UNIT, "cores": use, "kind": "reference",
            "sample": "%s; %d steps of one %d-sample block, %d host threads, SSE2 -O2 build" % (sample, steps, n, use),
            "ms_per_step": secs / steps * 1e3, "steps": steps, "rows": rows}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        r = reference_rate(args.workload, args.steps, args.warmup, full=True)
    except Exception as e:                                          # the oracle always exists; report why it did not run
        print(json.dumps({"impl": "reference", "unavailable": str(e)}))
        return
    dtype = WORKLOADS[args.workload][5]
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "impl": "reference",
            "config": workload_config(args.workload),
            "timing": {"blocks_per_step": 1, "note": "each step is a bounded sample of the workload: one block"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# synthetic data shared by the timed runs, the parity leg and tests/test_gpu_fullsize.py
# ------------------------------------------------------------------------------------------------------
def ir_seed(ins, outs, group, o, i):
    """seed of the impulse response of GLOBAL pair (group, in i, out o) (2000 + pair)"""
    return 2000 + (group * outs + o) * ins + i


def device_ir(gen, seed, taps, decay, tdt):
    """N(0,1) x exp(-6.9 k / L) on the device, reproducible from `seed` on any rank (same generator, same device type)"""
    import torch
    gen.manual_seed(seed)
    return torch.randn(taps, generator=gen, device=decay.device, dtype=tdt) * decay


def input_pool(gen, rank_seed, rows, n, n_pool, tdt, dev):
    """`n_pool` blocks of white noise uniform[-1, 1) for `rows` channels, reproducible from 1000 + rank_seed"""
    import torch
    gen.manual_seed(1000 + rank_seed)
    return [torch.rand(rows, n, generator=gen, device=dev, dtype=tdt) * 2 - 1 for _ in range(n_pool)]


def reference_rows(workload, irs_check, x_full):
    """The compiled reference on the checked rows: irs_check[row][in][taps] (float matrix) or [chan][taps] (double
    banks), x_full[in][n] or [chan][n].  Returns y[row][n] and the oracle's name."""
    import checkers as ck
    ins, outs, groups, taps, B, dtype, _ = WORKLOADS[workload]
    if dtype == "f64":
        y = np.stack([ck.ref_restated_run_f64(irs_check[c], x_full[c], 2 * B) for c in range(len(irs_check))])
        return y, "oracle/_ref ref_restated_f64: PartitionedConvolve.cpp:173-426 restated in double over the reference's own FFT_SETUP_D transforms"
    y = ck.ref_matrix_run(irs_check, x_full, 2 * B)
    return y, "oracle/_ref ref_matrix: rows of the unmodified MonoConvolve(maxLength, false, %d) summed as NToMonoConvolve.cpp:35-43" % (2 * B)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import hisstools_library_b200 as hb                              # noqa: F401
    from hisstools_library_b200 import _abi
    from hisstools_library_b200.convolve import _Engine
    import checkers as ck

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        # host-side barriers for the legs in which rank 0 alone drives every GPU (an NCCL barrier would park a spinning kernel on them)
        cpu_group = dist.new_group(backend="gloo")
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
    n = B
    P = (taps + B - 1) // B

    sharded = None
    if mode == "inputs":
        # the public multi-GPU class: inputs sharded over the ranks, partial output blocks summed across them
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
    if args.tail_streams:
        eng.set_tail_streams(args.tail_streams)
    # the input rows of every timed call sit complete in HBM before the call is made (the pool below), which is what mode 2 declares:
    # fused engines (configs 1-3) may then run consecutive single-block calls side by side.  Config 4 / 5 are not fused engines.
    eng.set_hop_overlap(args.hop_overlap)
    gen = torch.Generator(device=dev)
    decay = torch.exp(-6.9 * torch.arange(taps, device=dev, dtype=torch.float64) / taps).to(tdt)
    for g in range(l_groups):
        for o in range(outs):
            for i in range(l_ins):
                gi = (rank * l_ins + i) if mode == "inputs" else i
                gg = (rank * l_groups + g) if mode == "groups" else g
                ir = device_ir(gen, ir_seed(ins, outs, gg, o, gi), taps, decay, tdt)
                eng.set_ir_device(g, i, o, ir.data_ptr(), taps)
    torch.cuda.synchronize()
    assert eng.partitions == P, (eng.partitions, P)

    rows_in, rows_out = l_groups * l_ins, l_groups * outs
    n_pool = 4
    x_pool = input_pool(gen, rank, rows_in, n, n_pool, tdt, dev)
    y_part = torch.zeros(rows_out, n, device=dev, dtype=tdt)
    shard_rows = rows_out // world if mode == "inputs" else rows_out
    y_shard = torch.zeros(shard_rows, n, device=dev, dtype=tdt) if mode == "inputs" else y_part
    # a real (non-default) stream: handle 0 would mean "the engine's own stream" to the C ABI
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    # (pointers looked up once: a block of configs 1-3 takes a few microseconds, as long as a handful of Python attribute calls)
    x_ptrs = [t.data_ptr() for t in x_pool]
    y_ptr, y_ld, s_ptr = y_part.data_ptr(), y_part.stride(0), stream.cuda_stream

    def block(k, y=None):
        """one block of B samples per channel through the device-resident entry point"""
        if sharded is not None:
            sharded.process_device(x_pool[k % n_pool], y_shard if y is None else y, n, s_ptr)
        elif y is None:
            eng.process_device(x_ptrs[k % n_pool], n, y_ptr, y_ld, n, False, s_ptr)
        else:
            eng.process_device(x_ptrs[k % n_pool], n, y.data_ptr(), y.stride(0), n, False, s_ptr)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- calibration: blocks per step so that the timed region lasts >= ~0.6 s -----------------------
    for k in range(8):
        block(k)
    barrier()
    t_block = 0.0
    for n_cal in (16, 512, 4096):
        # (blocks of a few microseconds -- overlapping fused hops -- need more than 16 of them to show their steady rate)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for k in range(n_cal):
            block(k)
        eng.join(stream.cuda_stream)
        c1.record(stream)
        barrier()
        t_block = c0.elapsed_time(c1) / n_cal * 1e-3
        # every rank takes the same decision (the slowest rank's time), or their barriers would no longer pair up
        tc = torch.tensor([t_block], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        t_block = float(tc.item())
        if t_block * n_cal > 5e-3:
            break
    if args.blocks_per_step > 0:
        R = args.blocks_per_step
    else:
        tb = torch.tensor([t_block], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        R = int(min(65536, max(1, math.ceil(args.min_seconds / (max(args.steps, 1) * float(tb.item()))))))
    warm_steps = max(args.warmup, 3)

    k = 0
    for _ in range(warm_steps * R):
        block(k)
        k += 1
    barrier()

    # ---- device-resident timed region: K steps of R blocks, one call per block -------------------------
    lib = _abi.lib()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.hb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps * R):
        block(k)
        k += 1
    # the overlapped schedule keeps the tail of the NEXT block in flight on the engine's second stream: it is joined
    # before the closing event, so the window holds exactly as many tail launches as blocks (steady-state period)
    eng.join(stream.cuda_stream)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.hb_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # ---- second pass with CUDA events around every kernel launch (roofline figure); kept out of the timed region
    # above because the event records open small gaps between the kernels
    prof_blocks = min(args.steps * R, 128)
    tail_streams = eng.tail_streams
    if tail_streams == 2:
        # with two alternating tail streams a launch's events also span the time its CTAs wait for SMs: the kernel's own
        # duration is taken with one tail stream (the timed region above ran with two)
        eng.set_tail_streams(1)
        for _ in range(4):
            block(k)
            k += 1
    eng.set_profiling(True)
    for _ in range(prof_blocks):
        block(k)
        k += 1
    barrier()
    prof, hops = eng.get_profile()
    eng.set_profiling(False)
    if tail_streams == 2:
        eng.set_tail_streams(args.tail_streams)
        for _ in range(4):
            block(k)
            k += 1
        barrier()
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
    blocks = args.steps * R
    value = job_rows * n * blocks / (ms * 1e-3) / 1e6
    ms_per_block = ms / blocks
    if fused:
        # consecutive launches of the fused hop overlap (programmatic dependent launch: the next hop runs its partitions >= 2 beside
        # this one), so the events around one launch span more than a period; the period itself is the per-launch figure
        cmac_ms = ms_per_block

    # ---- multi-hop reuse: calls of 4 and 8 blocks, every IR spectrum streamed once per call (reported separately: the
    # per-hop byte figure above does not apply to it) --------------------------------------
    multi = None
    if not args.no_multi_hop:
        multi = []
        # HBM-bound engines: 4 / 8 blocks per call share one pass over the IR spectra; launch-latency-bound ones (configs 1-3): 8 / 64
        # blocks per call share their launches
        bytes_hint = eng.bytes_per_hop > 256e6
        for mh in ((4, 8) if bytes_hint else (8, 64)):
            xm = [torch.rand(rows_in, n * mh, generator=gen, device=dev, dtype=tdt) * 2 - 1 for _ in range(2)]
            ym = torch.zeros(shard_rows if sharded is not None else rows_out, n * mh, device=dev, dtype=tdt)

            def call(q):
                if sharded is not None:
                    sharded.process_device(xm[q % 2], ym, n * mh, stream.cuda_stream)
                else:
                    eng.process_device(xm[q % 2].data_ptr(), n * mh, ym.data_ptr(), n * mh, n * mh, False, stream.cuda_stream)
            for q in range(3):
                call(q)
            barrier()
            m_calls = int(max(6, min(4096, math.ceil(0.3 / max(t_block * 2, 1e-6)))))
            if world > 1:
                mc = torch.tensor([float(m_calls)], device=dev, dtype=torch.float64)
                dist.all_reduce(mc, op=dist.ReduceOp.MAX)
                m_calls = int(mc.item())
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record(stream)
            for q in range(m_calls):
                call(q)
            eng.join(stream.cuda_stream)
            m1.record(stream)
            barrier()
            mt = torch.tensor([m0.elapsed_time(m1) / m_calls], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(mt, op=dist.ReduceOp.MAX)
            mms = float(mt.item())
            multi.append({"blocks_per_call": mh, "value": job_rows * n * mh / (mms * 1e-3) / 1e6, "unit": UNIT, "ms_per_call": mms,
                          "ms_per_block": mms / mh, "calls": m_calls})
            del xm, ym
        # leave the engine as the single-block path found it (a tail in flight for the next block)
        block(k)
        k += 1
        barrier()

    # ---- end to end through the host-pointer boundary ----------------------------------------------
    e2e_blocks = max(3, args.steps) * R
    e2e_warm = max(3, args.warmup) * min(R, 8)
    h2d = rows_in * n * es * R
    if world == 1:
        xin = [np.ascontiguousarray(x_pool[q].cpu().numpy()) for q in range(n_pool)]
        yout = np.zeros((rows_out, n), ndt)
        rows_y = [yout[r] for r in range(rows_out)]
        rows_x = [[xi[r] for r in range(rows_in)] for xi in xin]
        # (the row-pointer arrays of these reused rows are built once, as a C++ caller's are: numpy's .ctypes costs about a
        # microsecond per row, 35 us per call at config 5 against a hop of 87)
        ptr_x = [eng.row_pointers(rx) for rx in rows_x]
        ptr_y = eng.row_pointers(rows_y)
        for q in range(e2e_warm):                                                         # staging buffers, copy streams, events
            eng.process_pointers(ptr_x[q % n_pool], ptr_y, n)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for q in range(e2e_blocks):
            # hb_conv_process: gathers the host rows, H2D, the block's kernels, D2H, scatters the block to the host rows
            eng.process_pointers(ptr_x[q % n_pool], ptr_y, n)
        torch.cuda.synchronize()                                                          # the last call's device work is inside the timed region
        e2e_s = time.perf_counter() - t0
        d2h = rows_out * n * es * R
        e2e_api = "hb_conv_process (host pointers), one call per block"
    else:
        xh = [x_pool[q].cpu().pin_memory() for q in range(n_pool)]
        yh = torch.zeros(y_shard.shape, dtype=tdt).pin_memory()
        xd = torch.empty_like(x_pool[0])
        got = torch.cuda.Event()

        def e2e_block(q):
            xd.copy_(xh[q % n_pool], non_blocking=True)
            if sharded is not None:
                sharded.process_device(xd, y_shard, n, stream.cuda_stream)
            else:
                eng.process_device(xd.data_ptr(), n, y_part.data_ptr(), n, n, False, stream.cuda_stream)
            yh.copy_(y_shard, non_blocking=True)
            # the block is over when its result is in host memory; the tail of the NEXT block, launched ahead on the engine's
            # second stream, keeps running behind this wait as it does in the device-resident loop (the final barrier +
            # synchronize below puts what is left of it inside the timed region)
            got.record(stream)
            got.synchronize()
        for q in range(e2e_warm):
            e2e_block(q)
        barrier()
        t0 = time.perf_counter()
        for q in range(e2e_blocks):
            e2e_block(q)
        torch.cuda.synchronize()
        barrier()
        e2e_s = time.perf_counter() - t0
        d2h = y_shard.numel() * es * R
        tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
        e2e_api = "per rank: pinned H2D + ShardedConvolver.process_device (%s) + D2H of this rank's output rows, waiting on the result event" % \
            ("hb_conv_process_shard_dev: peer stores from the inverse-FFT epilogue" if sharded is not None and sharded.exchange == "fused"
             else "hb_matrix_process_dev + NCCL reduce_scatter")
    e2e_value = job_rows * n * e2e_blocks / e2e_s / 1e6
    e2e_ms_per_block = e2e_s / e2e_blocks * 1e3

    # ---- parity against the compiled reference, outside every timed region ---------------------------
    parity, ref_pack = None, None
    if not args.no_parity:
        parity, ref_pack = parity_leg(args, torch, dist, ck, eng, sharded, gen, decay, stream, dev, mode, world, rank, l_ins, l_groups, tdt, ndt, n_pool)

    # ---- N > 1: end to end through ONE object in ONE process -- the reference's Convolver::process(ins, outs, ...) with host
    # pointers over all N GPUs (hb_matrix_create_multi); rank 0 drives every GPU, the other ranks wait on the host
    per_rank_e2e = None
    if world > 1 and mode in ("inputs", "groups") and not args.no_front:
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)
        front = None
        if rank == 0:
            front = front_leg(args, torch, ck, world, local, e2e_blocks, max(3, args.warmup) * min(R, 8), n_pool, tdt, ndt, ref_pack)
        dist.barrier(group=cpu_group)
        torch.cuda.set_device(local)
        if rank == 0 and front is not None:
            per_rank_e2e = {"value": e2e_value, "ms_per_block": e2e_ms_per_block, "api": e2e_api}
            e2e_value, e2e_ms_per_block, e2e_api = front["value"], front["ms_per_block"], front["api"]
            h2d, d2h = front["h2d"] * R, front["d2h"] * R
            if parity is not None and front.get("rel_rms") is not None:
                parity["rel_rms_host_pointer_path"] = front["rel_rms"]
                parity["host_pointer_path"] = front["api"]
                parity["ok"] = bool(parity["ok"] and front["rel_rms"] <= TOL[dtype])

    # ---- roofline of the dominant kernel (multiply-accumulate), per rank ---------------------------
    peak, peak_src = measured_peak()
    bytes_per_hop = eng.bytes_per_hop                             # IR spectra + FDL + time-domain I/O of one hop
    bytes_per_launch = eng.bytes_per_launch                       # the dominant launch's share of it
    achieved = bytes_per_launch / (cmac_ms * 1e-3) / 1e9 if cmac_ms > 0 else 0.0
    traffic, traffic_src = committed_traffic(args.workload, world, overlapped)
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src,
            "kernel": "k_cmac, tail launch: partitions 1..P-1 (frequency-domain multiply-accumulate)" if overlapped
                      else ("k_hop_fused (whole hop in one cluster launch)" if fused else "k_cmac (frequency-domain multiply-accumulate, all partitions)"),
            "bytes_per_launch": bytes_per_launch, "bytes_per_hop": bytes_per_hop, "kernel_ms": cmac_ms, "forward_fft_ms": fwd_ms,
            "head_cmac_ms": head_ms, ("wait_for_tail_plus_inverse_fft_ms" if overlapped else "inverse_fft_ms"): inv_ms,
            "kernel_share_of_step": cmac_ms / ms_per_block, "peak_source": peak_src,
            "kernel_ms_note": ("launches overlap (programmatic dependent launch): kernel_ms is the hop period" if fused else
                               ("measured with ONE tail stream; the timed region ran with two alternating tail streams, whose launches overlap "
                                "(ramp-up and drain hidden), so a launch's own duration may exceed the hop period" if tail_streams == 2 else
                                "CUDA events on the tail stream around the launch, in situ")),
            "hop_frac": bytes_per_hop / (ms_per_block * 1e-3) / 1e9 / peak}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            r = reference_rate(args.workload, 0, 0, budget_s=args.cpu_seconds)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm_steps,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if mode == "replicas" else "strong",
                "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": workload_config(args.workload),
                "timing": {"blocks_per_step": R, "ms_per_block": ms_per_block, "timed_region_s": ms * 1e-3,
                           "calls": "one process call per block of %d samples; a step is %d consecutive blocks" % (B, R),
                           "window": "CUDA events on the launching stream; the engine's look-ahead (tail) stream is joined before the closing event"},
                "engine": {"sharding": mode, "local_inputs": l_ins, "outputs": outs, "groups": l_groups, "schedule": eng.schedule,
                           "tail_streams": tail_streams,
                           "hop_overlap": ("consecutive single-block calls overlap (hb_conv_set_hop_overlap mode %d: input rows complete before each call)" % args.hop_overlap
                                           if args.hop_overlap else "every hop behind the previous one") if eng.schedule == "fused" else "n/a (not a fused engine)",
                           "transforms": {1: "one CTA each", 2: "cluster of 8 CTAs each (DSMEM)", 3: "four-step chains"}.get(eng.fft_path, "?"),
                           "l2": "inputs larger than L2: %.2f GiB of IR spectra per rank streamed every block" % (bytes_per_hop / 2 ** 30)
                                 if bytes_per_hop > 256e6 else "working set %.1f MiB is L2-resident (not an HBM-roofline case)" % (bytes_per_hop / 2 ** 20),
                           "collective": ("peer stores fused into the inverse-FFT epilogue (NVLink), owner-side sum" if sharded is not None and sharded.exchange == "fused"
                                          else "nccl reduce_scatter of partial output blocks") if mode == "inputs" else "none"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "blocks": e2e_blocks,
                        "ms_per_block": e2e_ms_per_block, "api": e2e_api},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof}
        if per_rank_e2e is not None:
            line["e2e"]["per_rank_processes"] = per_rank_e2e
        if parity is not None:
            line["parity"] = parity
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if multi is not None:
            line["multi_hop_reuse"] = {"runs": multi, "note": "hop-aligned calls of several blocks: on HBM-bound engines every IR spectrum is read "
                                                              "once per call (k_cmac_mh2); on launch-latency-bound engines the blocks of a call share their launches"}
        print(json.dumps(line))
    if sharded is not None:
        sharded.close()
    else:
        eng.close()
    if world > 1:
        dist.destroy_process_group()


def parity_leg(args, torch, dist, ck, eng, sharded, gen, decay, stream, dev, mode, world, rank, l_ins, l_groups, tdt, ndt, n_pool):
    """P + 16 blocks of the synthetic stream through the benchmarked engine (device-resident calls, and at N = 1 the
    host-pointer calls as well) after a reset, against the compiled reference on the rows rank 0 owns."""
    ins, outs, groups, taps, B, dtype, _ = WORKLOADS[args.workload]
    P = (taps + B - 1) // B
    hops = P + 16
    n = B
    rows_in = l_groups * l_ins
    if ck.ref() is None:
        return {"rel_rms": None, "oracle": "unavailable: oracle/_ref/libhisstools_ref.so is missing"}, None
    # the rows of this rank's output that are checked: first and last it owns
    if mode == "inputs":
        own = outs // world
        check_local = sorted(set([0, own - 1]))
        check_global = check_local                                   # rank 0 owns outputs 0 .. own-1
    elif groups > 1:
        check_local = sorted(set([0, l_groups - 1]))                 # independent banks: bank index = row index
        check_global = check_local
    else:
        check_local = sorted(set([0, outs - 1]))
        check_global = check_local
    eng.reset()
    x_pool = input_pool(gen, rank, rows_in, n, n_pool, tdt, dev)
    y_all = torch.zeros(len(check_local), hops * n, device=dev, dtype=tdt)
    shard_rows = outs // world if mode == "inputs" else l_groups * outs
    yb = torch.zeros(shard_rows, n, device=dev, dtype=tdt)
    for k in range(hops):
        x = x_pool[k % n_pool]
        if sharded is not None:
            sharded.process_device(x, yb, n, stream.cuda_stream)
        else:
            eng.process_device(x.data_ptr(), n, yb.data_ptr(), n, n, False, stream.cuda_stream)
        for q, r in enumerate(check_local):
            y_all[q, k * n:(k + 1) * n].copy_(yb[r], non_blocking=True)
    torch.cuda.synchronize()
    got = y_all.cpu().numpy()
    got_host = None
    if world == 1:
        # the same stream through the host-pointer entry point (the e2e path)
        eng.reset()
        xin = [np.ascontiguousarray(x_pool[q].cpu().numpy()) for q in range(n_pool)]
        yout = np.zeros((shard_rows, n), ndt)
        rows_y = [yout[r] for r in range(shard_rows)]
        got_host = np.zeros((len(check_local), hops * n), ndt)
        for k in range(hops):
            eng.process([xin[k % n_pool][r] for r in range(rows_in)], rows_y, n)
            for q, r in enumerate(check_local):
                got_host[q, k * n:(k + 1) * n] = yout[r]
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank != 0:
        return None, None
    # the reference's inputs on rank 0: every rank's pool regenerated from its seed, the checked rows' IRs from theirs
    if mode in ("inputs", "groups"):
        # the whole job's rows, rank after rank (rank 0's come first)
        pools = [input_pool(gen, r, rows_in, n, n_pool, tdt, dev) for r in range(world)]
        x_full = np.concatenate([np.concatenate([pools[r][k % n_pool].cpu().numpy() for k in range(hops)], axis=1) for r in range(world)], axis=0)
        del pools
    else:
        x_full = np.concatenate([x_pool[k % n_pool].cpu().numpy() for k in range(hops)], axis=1)
    if groups > 1:
        irs_check = np.stack([device_ir(gen, ir_seed(ins, outs, g, 0, 0), taps, decay, tdt).cpu().numpy() for g in check_global])
        x_ref = x_full[check_local]
    else:
        irs_check = np.stack([np.stack([device_ir(gen, ir_seed(ins, outs, 0, o, i), taps, decay, tdt).cpu().numpy() for i in range(ins)])
                              for o in check_global])
        x_ref = x_full
    t0 = time.perf_counter()
    want, oracle = reference_rows(args.workload, irs_check, x_ref)
    ref_s = time.perf_counter() - t0
    worst = max(ck.rel_rms(got[q], want[q]) for q in range(len(check_local)))
    last = max(ck.rel_rms(got[q][-16 * n:], want[q][-16 * n:]) for q in range(len(check_local)))
    res = {"rel_rms": worst, "rel_rms_last_16_blocks": last, "tolerance": TOL[dtype], "ok": bool(worst <= TOL[dtype] and last <= TOL[dtype]),
           "rows": ["bank %d" % g for g in check_global] if groups > 1 else ["output %d x all %d inputs" % (o, ins) for o in check_global],
           "blocks": hops, "path": "device-resident calls, one block each (the timed path)" + (", %d ranks" % world if world > 1 else ""),
           "oracle": oracle, "oracle_seconds": ref_s}
    if got_host is not None:
        res["rel_rms_host_pointer_path"] = max(ck.rel_rms(got_host[q], want[q]) for q in range(len(check_local)))
        res["ok"] = bool(res["ok"] and res["rel_rms_host_pointer_path"] <= TOL[dtype])
    # what the single-process multi-GPU leg needs to repeat the check: the whole job's input stream, the checked rows, the reference
    x_job = x_full
    return res, {"x": x_job, "rows": check_global, "want": want, "hops": hops}


def front_leg(args, torch, ck, world, local, blocks, warm, n_pool, tdt, ndt, ref_pack):
    """Rank 0 only: the whole workload behind ONE hb_matrix handle dealt to all `world` GPUs of this process, driven through
    the host-pointer C-ABI call (what HISSTools::Convolver::process forwards to), one call per block: H2D, the kernels of every
    device, the exchange, D2H inside the timed region.  Also streams the parity input through it when the reference rows exist."""
    from hisstools_library_b200 import _abi
    from hisstools_library_b200.convolve import _Matrix
    ins, outs, groups, taps, B, dtype, _ = WORKLOADS[args.workload]
    es = 8 if dtype == "f64" else 4
    n = B
    devices = list(range(world))
    m = _Matrix(groups, ins, outs, taps, (False, 2 * B, 0, 0, 0), ndt, 0, devices)
    m.setResetOffset(0)
    engines = m.shard_engines()
    by_inputs = groups == 1
    l_ins, l_groups = (ins // world, 1) if by_inputs else (ins, groups // world)
    for d in devices:
        dd = torch.device("cuda", d)
        with torch.cuda.device(dd):
            gen = torch.Generator(device=dd)
            decay = torch.exp(-6.9 * torch.arange(taps, device=dd, dtype=torch.float64) / taps).to(tdt)
            for g in range(l_groups):
                for o in range(outs):
                    for i in range(l_ins):
                        gi = d * l_ins + i if by_inputs else i
                        gg = g if by_inputs else d * l_groups + g
                        ir = device_ir(gen, ir_seed(ins, outs, gg, o, gi), taps, decay, tdt)
                        engines[d].set_ir_device(g, i, o, ir.data_ptr(), taps)
            torch.cuda.synchronize()
    torch.cuda.set_device(local)
    rows_in, rows_out = groups * ins, groups * outs
    rng = np.random.default_rng(11)
    xin = [rng.uniform(-1, 1, (rows_in, n)).astype(ndt) for _ in range(n_pool)]
    yout = np.zeros((rows_out, n), ndt)
    lib = _abi.lib()
    VP = C.c_void_p
    ips = [(VP * rows_in)(*[x[r].ctypes.data for r in range(rows_in)]) for x in xin]
    ops = (VP * rows_out)(*[yout[r].ctypes.data for r in range(rows_out)])
    h = m._h
    for q in range(warm):
        _abi.check(lib.hb_matrix_process(h, ips[q % n_pool], ops, n, 0))
    for d in devices:
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for q in range(blocks):
        _abi.check(lib.hb_matrix_process(h, ips[q % n_pool], ops, n, 0))
    for d in devices:
        torch.cuda.synchronize(d)                                       # the last call's device work is inside the timed region
    secs = time.perf_counter() - t0
    res = {"value": rows_out * n * blocks / secs / 1e6, "ms_per_block": secs / blocks * 1e3, "h2d": rows_in * n * es, "d2h": rows_out * n * es,
           "api": "hb_matrix_process (host pointers) on ONE multi-device matrix (hb_matrix_create_multi, %d GPUs, exchange: %s), one call per block "
                  "from one process" % (world, m.exchange), "rel_rms": None}
    if ref_pack is not None:
        m.reset()
        x = ref_pack["x"]
        hops = ref_pack["hops"]
        got = np.zeros((len(ref_pack["rows"]), hops * n), ndt)
        xb = np.zeros((rows_in, n), ndt)
        ip = (VP * rows_in)(*[xb[r].ctypes.data for r in range(rows_in)])
        for k in range(hops):
            xb[:] = x[:, k * n:(k + 1) * n]
            _abi.check(lib.hb_matrix_process(h, ip, ops, n, 0))
            for q, r in enumerate(ref_pack["rows"]):
                got[q, k * n:(k + 1) * n] = yout[r]
        res["rel_rms"] = max(ck.rel_rms(got[q], ref_pack["want"][q]) for q in range(len(ref_pack["rows"])))
    m.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--blocks-per-step", type=int, default=0, help="blocks of B samples per step, one call each (0: as many as make the timed region last --min-seconds)")
    ap.add_argument("--min-seconds", type=float, default=0.6, help="shortest timed region the automatic blocks-per-step aims at")
    ap.add_argument("--variant", type=int, default=None, help="multiply-accumulate kernel: 1 = TMA ring, 0 = direct loads")
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--schedule", default=None, choices=["overlapped", "serial"], help="hop schedule (default: the library's automatic choice)")
    ap.add_argument("--fft-path", type=int, default=0, choices=[0, 1, 2, 3],
                    help="transforms: 0 automatic, 1 one CTA each, 2 cluster of 8 CTAs each, 3 four-step chains (hb_conv_set_fft_path)")
    ap.add_argument("--tail-streams", type=int, default=0, choices=[0, 1, 2], help="overlapped schedule: streams the tail launches alternate between (0: library default)")
    ap.add_argument("--hop-overlap", type=int, default=2, choices=[0, 1, 2],
                    help="fused engines (configs 1-3): 0 every hop behind the previous one, 2 consecutive calls overlap (hb_conv_set_hop_overlap)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "fused", "nccl"], help="multi-GPU sum of partial outputs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-multi-hop", action="store_true", help="skip the multi-hop reuse leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity leg")
    ap.add_argument("--no-front", action="store_true", help="N > 1: skip the single-process multi-GPU end-to-end leg (report the per-rank one)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
