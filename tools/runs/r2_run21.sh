#!/bin/bash
# round 2: end-to-end leg with the row-pointer arrays built once (as a C++ caller's are); direct pinned input on / off for configs 1-3
mkdir -p gpurun_out
for d in 1 0; do
for wl in c1 c2 c3; do
  HB_NO_DIRECT_IN=$d timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-multi-hop --no-parity > gpurun_out/r2ai_bench_${wl}_nd$d.json 2> gpurun_out/r2ai_bench_${wl}_nd$d.err
done
done
for wl in c5 c4; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-multi-hop --no-parity > gpurun_out/r2ai_bench_${wl}.json 2> gpurun_out/r2ai_bench_${wl}.err
done
python - <<'PY'
import json
for d in (1,0):
  for wl in ('c1','c2','c3'):
    try:
        x=json.loads(open('gpurun_out/r2ai_bench_%s_nd%d.json'%(wl,d)).read().strip().splitlines()[-1])
        print('%s direct_in %s value %.1f e2e %.1f' % (wl, 'off' if d else 'on', x['value'], x['e2e']['value']))
    except Exception as e: print(wl,d,'failed',e)
for wl in ('c5','c4'):
    x=json.loads(open('gpurun_out/r2ai_bench_%s.json'%wl).read().strip().splitlines()[-1])
    print('%s value %.1f e2e %.1f' % (wl, x['value'], x['e2e']['value']))
PY
