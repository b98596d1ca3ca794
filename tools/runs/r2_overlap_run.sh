#!/bin/bash
# round 2: consecutive fused hops that overlap (hb_conv_set_hop_overlap) -- tests, then configs 1-3 with and without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "fused or hop_batches or call_sizes or golden or semantics" > gpurun_out/r2t_pytest.log 2>&1
tail -5 gpurun_out/r2t_pytest.log
for ov in 0 2; do
for wl in c1 c2 c3; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --cpu-seconds 3 --hop-overlap $ov > gpurun_out/r2t_bench_${wl}_ov$ov.json 2> gpurun_out/r2t_bench_${wl}_ov$ov.err
done
done
python - <<'PY'
import json
for ov in (0, 2):
  for wl in ('c1','c2','c3'):
    try:
        d=json.loads(open('gpurun_out/r2t_bench_%s_ov%d.json'%(wl,ov)).read().strip().splitlines()[-1])
        print('%s ov %d value %.1f e2e %.1f us/block %.2f multi %s parity %.2e/%s cpu %.1f clk %s' % (wl, ov, d['value'], d['e2e']['value'], d['timing']['ms_per_block']*1e3, [(m['blocks_per_call'], round(m['value'],1)) for m in d['multi_hop_reuse']['runs']], d['parity']['rel_rms'], d['parity']['ok'], d['cpu_baseline']['value'], d['clocks']['sm_mhz']))
    except Exception as e: print(wl,ov,'failed',e)
PY
