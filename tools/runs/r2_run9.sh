#!/bin/bash
mkdir -p gpurun_out
python tools/fft_family_bench.py > gpurun_out/r2n_fft_family.txt 2>&1; cat gpurun_out/r2n_fft_family.txt | tail -30
python -m pytest tests/test_gpu_spectral.py -x -q -k kernel_smoother 2>&1 | tail -3
for cfg in "8 256" "8 2200" "16 1100" "16 2200"; do
  set -- $cfg
  HB_FUSED_MAX_CS=$1 HB_FUSED_MAX_KB=$2 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu --no-multi-hop > gpurun_out/r2n_c3_$1_$2.json 2> gpurun_out/r2n_c3_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2n_c3_$1_$2.json').read().strip().splitlines()[-1])
    print('c3 max_cs=$1 max_kb=$2 schedule %s value %.1f us/block %.2f e2e %.1f parity %s' % (d['engine']['schedule'], d['value'], d['timing']['ms_per_block']*1e3, d['e2e']['value'], d['parity']['rel_rms']))
except Exception as e: print('failed', e)
PY
  tail -2 gpurun_out/r2n_c3_$1_$2.err
done
