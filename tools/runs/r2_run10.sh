#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2o_pytest.log; tail -30 gpurun_out/r2o_pytest.log
python tools/fft_family_bench.py > gpurun_out/r2o_fft_family.txt 2>&1; tail -26 gpurun_out/r2o_fft_family.txt
python bench.py --workload c3 --steps 20 --warmup 5 --cpu-seconds 3 > gpurun_out/r2o_bench_c3.json 2> gpurun_out/r2o_bench_c3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench_c3.json').read().strip().splitlines()[-1])
print('c3 value %.1f us/block %.2f e2e %.1f multi %s parity %s sched %s' % (d['value'], d['timing']['ms_per_block']*1e3, d['e2e']['value'], [(m['blocks_per_call'], round(m['value'],1)) for m in d['multi_hop_reuse']['runs']], d['parity']['ok'], d['engine']['schedule']))
PY
