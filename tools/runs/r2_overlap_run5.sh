#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_fullsize.py -m gpu -x -q -k "fused or overlap or bench_parity or call_sizes or golden" > gpurun_out/r2y_pytest.log 2>&1; tail -30 gpurun_out/r2y_pytest.log | grep -v "^$" | tail -12
{
for wl in c1 c2 c3; do for m in 0 2; do timeout 120 python tools/fused_chain_trace.py $wl $m; done; done
} > gpurun_out/r2y_chain_trace.txt 2>&1
cat gpurun_out/r2y_chain_trace.txt
nvcc -O2 -o /tmp/hop_rate tools/hop_rate.cu -Iinclude -Lhisstools_library_b200/lib -lhisstools_b200 -Xlinker -rpath=$PWD/hisstools_library_b200/lib -Wno-deprecated-gpu-targets \
  && timeout 300 /tmp/hop_rate > gpurun_out/r2y_hop_rate.txt 2>&1
cat gpurun_out/r2y_hop_rate.txt
for wl in c1 c2 c3; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --cpu-seconds 3 > gpurun_out/r2y_bench_${wl}.json 2> gpurun_out/r2y_bench_${wl}.err
done
python - <<'PY'
import json
for wl in ('c1','c2','c3'):
    try:
        d=json.loads(open('gpurun_out/r2y_bench_%s.json'%wl).read().strip().splitlines()[-1])
        print('%s value %.1f e2e %.1f us/block %.2f multi %s parity %.2e/%s cpu %.1f clk %s' % (wl, d['value'], d['e2e']['value'], d['timing']['ms_per_block']*1e3, [(m['blocks_per_call'], round(m['value'],1)) for m in d['multi_hop_reuse']['runs']], d['parity']['rel_rms'], d['parity']['ok'], d['cpu_baseline']['value'], d['clocks']['sm_mhz']))
    except Exception as e: print(wl,'failed',e)
PY
