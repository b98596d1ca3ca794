#!/bin/bash
# round 2: host-pointer path with rows moved in groups -- tests, config 5 / config 4 end to end
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_fullsize.py tests/test_gpu_cpp_dropin.py -m gpu -x -q > gpurun_out/r2ag_pytest.log 2>&1; tail -4 gpurun_out/r2ag_pytest.log
for wl in c5 c4; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-multi-hop --no-parity > gpurun_out/r2ag_bench_${wl}.json 2> gpurun_out/r2ag_bench_${wl}.err
done
python - <<'PY'
import json
for wl in ('c5','c4'):
    try:
        d=json.loads(open('gpurun_out/r2ag_bench_%s.json'%wl).read().strip().splitlines()[-1])
        print('%s value %.1f e2e %.1f' % (wl, d['value'], d['e2e']['value']))
    except Exception as e: print(wl,'failed',e)
PY
