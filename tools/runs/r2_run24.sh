#!/bin/bash
# round 2: host-pointer path with the finished block packed on the device (one contiguous download) -- tests, config 5 / 4 end to end, phases
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_fullsize.py tests/test_gpu_cpp_dropin.py -m gpu -x -q > gpurun_out/r2ao_pytest.log 2>&1; tail -3 gpurun_out/r2ao_pytest.log
HB_HOST_TRACE=1 timeout 300 python tools/diag_c5_e2e.py trace 2>&1 | tail -4
for wl in c5 c4; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-multi-hop > gpurun_out/r2ao_bench_${wl}.json 2> gpurun_out/r2ao_bench_${wl}.err
done
python - <<'PY'
import json
for wl in ('c5','c4'):
    try:
        d=json.loads(open('gpurun_out/r2ao_bench_%s.json'%wl).read().strip().splitlines()[-1])
        print('%s value %.1f e2e %.1f parity %.2e/%s' % (wl, d['value'], d['e2e']['value'], d['parity']['rel_rms'], d['parity']['ok']))
    except Exception as e: print(wl,'failed',e)
PY
