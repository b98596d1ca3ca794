"""Development probe (GPU box): per-hop time of the engine at a given shape for both k_cmac variants.
usage: python tools/gpu_probe.py [ins outs B P groups dtype]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hisstools_library_b200 as hb
from hisstools_library_b200.convolve import _Engine

def run(ins, outs, B, P, groups=1, dtype="f32", hops=12, variants=(1, 0), cps=(0,)):
    tdt = torch.float32 if dtype == "f32" else torch.float64
    ndt = np.float32 if dtype == "f32" else np.float64
    L = B * P
    t0 = time.time()
    e = _Engine(ndt, groups, ins, outs, 2 * B, L, 0, 0)
    e.set_reset_offset(0)
    g = torch.Generator(device="cuda").manual_seed(1)
    decay = torch.exp(-6.9 * torch.arange(L, device="cuda", dtype=tdt) / L)
    for gr in range(groups):
        for o in range(outs):
            for i in range(ins):
                ir = torch.randn(L, device="cuda", dtype=tdt, generator=g) * decay
                e.set_ir_device(gr, i, o, ir.data_ptr(), L)
    torch.cuda.synchronize()
    print("setup %.1fs  bytes/hop %.1f MB" % (time.time() - t0, e.bytes_per_hop / 1e6), flush=True)
    x = torch.rand(groups * ins, B, device="cuda", dtype=tdt) * 2 - 1
    y = torch.zeros(groups * outs, B, device="cuda", dtype=tdt)
    st = torch.cuda.current_stream().cuda_stream
    for variant in variants:
        for c in cps:
            e.set_tuning(c, variant)
            for _ in range(3):
                e.process_device(x.data_ptr(), B, y.data_ptr(), B, B, False, st)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(hops + 1)]
            ev[0].record()
            for h in range(hops):
                e.process_device(x.data_ptr(), B, y.data_ptr(), B, B, False, st)
                ev[h + 1].record()
            torch.cuda.synchronize()
            ts = [ev[h].elapsed_time(ev[h + 1]) for h in range(hops)]
            ms = float(np.median(ts))
            gbs = e.bytes_per_hop / ms / 1e6
            print("variant %d cps %d: %.1f us/hop (min %.1f max %.1f)  %.0f GB/s  %.3f of 6549  | %.1f Msamples/s  y.rms %.4g"
                  % (variant, c, ms * 1e3, min(ts) * 1e3, max(ts) * 1e3, gbs, gbs / 6549.1, groups * outs * B / ms / 1e3, float(y.double().pow(2).mean().sqrt())), flush=True)
    e.close()

if __name__ == "__main__":
    a = sys.argv[1:]
    if a:
        run(int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]) if len(a) > 4 else 1, a[5] if len(a) > 5 else "f32")
    else:
        run(8, 64, 4096, 64)                # config 4, one GPU's share when sharded over 8
        run(64, 64, 4096, 64)               # config 4 on one GPU
        run(1, 1, 8192, 128, 16, "f64")     # config 5
        run(8, 1, 2048, 64)                 # config 3
        run(1, 1, 1024, 64)                 # config 2
