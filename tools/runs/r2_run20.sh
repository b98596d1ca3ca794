#!/bin/bash
# round 2: host-pointer path of small fused engines reading their rows from pinned memory -- tests, configs 1-3 end to end with and without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_cpp_dropin.py tests/test_gpu_spectral.py -m gpu -x -q > gpurun_out/r2ah_pytest.log 2>&1; tail -3 gpurun_out/r2ah_pytest.log
for d in 1 0; do
for wl in c1 c2 c3; do
  HB_NO_DIRECT_IN=$d timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-multi-hop --no-parity > gpurun_out/r2ah_bench_${wl}_nd$d.json 2> gpurun_out/r2ah_bench_${wl}_nd$d.err
done
done
python - <<'PY'
import json
for d in (1,0):
  for wl in ('c1','c2','c3'):
    try:
        x=json.loads(open('gpurun_out/r2ah_bench_%s_nd%d.json'%(wl,d)).read().strip().splitlines()[-1])
        print('%s direct_in %s value %.1f e2e %.1f' % (wl, 'off' if d else 'on', x['value'], x['e2e']['value']))
    except Exception as e: print(wl,d,'failed',e)
PY
