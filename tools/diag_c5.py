"""per-channel parity of the config-5 engine against the restated double reference, for several call patterns / FFT paths"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import checkers as ck, bench
from hisstools_library_b200.convolve import _Engine

P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
groups, B = 16, 8192
taps, hops = B * P, P + 8
dev = torch.device("cuda", 0)
for path in (0, 1):
    eng = _Engine(np.float64, groups, 1, 1, 2 * B, taps, 0, 0, 0)
    eng.set_reset_offset(0)
    if path:
        eng.set_fft_path(path)
    gen = torch.Generator(device=dev)
    decay = torch.exp(-6.9 * torch.arange(taps, device=dev, dtype=torch.float64) / taps)
    irs = []
    for g in range(groups):
        ir = bench.device_ir(gen, bench.ir_seed(1, 1, g, 0, 0), taps, decay, torch.float64)
        eng.set_ir_device(g, 0, 0, ir.data_ptr(), taps)
        irs.append(ir.cpu().numpy())
    pool = bench.input_pool(gen, 0, groups, B, 4, torch.float64, dev)
    xs = torch.cat([pool[k % 4] for k in range(hops)], dim=1).contiguous()
    xh = xs.cpu().numpy()
    want = np.stack([ck.ref_restated_run_f64(irs[g], xh[g], 2 * B) for g in range(groups)])
    for calls in ([1], [4], [1, 2, 5]):
        eng.reset()
        n = xs.shape[1]
        y = torch.zeros(groups, n, device=dev, dtype=torch.float64)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            pos, k = 0, 0
            while pos < n:
                m = min(calls[k % len(calls)] * B, n - pos)
                eng.process_device(xs[:, pos:].data_ptr(), xs.stride(0), y[:, pos:].data_ptr(), y.stride(0), m, False, st.cuda_stream)
                pos += m; k += 1
        torch.cuda.synchronize()
        got = y.cpu().numpy()
        errs = [ck.rel_rms(got[g], want[g]) for g in range(groups)]
        print("path", path, "fft_path", eng.fft_path, "calls", calls, " ".join("%.1e" % e for e in errs), flush=True)
        bad = [g for g in range(groups) if errs[g] > 1e-12]
        for g in bad[:2]:
            per_hop = [ck.rel_rms(got[g][h * B:(h + 1) * B], want[g][h * B:(h + 1) * B]) for h in range(hops)]
            print("   ch", g, "per-hop:", " ".join("%.0e" % e for e in per_hop[:40]), flush=True)
    eng.close()
