#!/bin/bash
# round 2: programmatic dependent launch of the fused hop -- parity, then configs 1 / 2 with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_cpp_dropin.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -6
for wl in c1 c2; do
  for pdl in 0 1; do
    HB_NO_PDL=$pdl python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-multi-hop > gpurun_out/r2m_${wl}_nopdl$pdl.json 2> gpurun_out/r2m_${wl}_nopdl$pdl.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/r2m_${wl}_nopdl$pdl.json').read().strip().splitlines()[-1])
print('$wl HB_NO_PDL=$pdl value %.1f us/block %.2f e2e %.1f parity %s %s' % (d['value'], d['timing']['ms_per_block']*1e3, d['e2e']['value'], d['parity']['rel_rms'], d['parity']['ok']))
PY
    tail -2 gpurun_out/r2m_${wl}_nopdl$pdl.err
  done
done
