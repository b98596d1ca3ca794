#!/bin/bash
# round 2: overlapping fused hops -- whole 1-GPU suite, the C++ call-rate tool, configs 1-3
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1
tail -4 gpurun_out/r2u_pytest.log
nvcc -O2 -o /tmp/hop_rate tools/hop_rate.cu -Iinclude -Lhisstools_library_b200/lib -lhisstools_b200 -Xlinker -rpath=$PWD/hisstools_library_b200/lib -Wno-deprecated-gpu-targets \
  && timeout 300 /tmp/hop_rate > gpurun_out/r2u_hop_rate.txt 2>&1
cat gpurun_out/r2u_hop_rate.txt
for wl in c1 c2 c3; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --cpu-seconds 3 > gpurun_out/r2u_bench_${wl}.json 2> gpurun_out/r2u_bench_${wl}.err
done
python - <<'PY'
import json
for wl in ('c1','c2','c3'):
    try:
        d=json.loads(open('gpurun_out/r2u_bench_%s.json'%wl).read().strip().splitlines()[-1])
        print('%s value %.1f e2e %.1f us/block %.2f multi %s parity %.2e/%s cpu %.1f clk %s' % (wl, d['value'], d['e2e']['value'], d['timing']['ms_per_block']*1e3, [(m['blocks_per_call'], round(m['value'],1)) for m in d['multi_hop_reuse']['runs']], d['parity']['rel_rms'], d['parity']['ok'], d['cpu_baseline']['value'], d['clocks']['sm_mhz']))
    except Exception as e: print(wl,'failed',e)
PY
