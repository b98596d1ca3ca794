#!/bin/bash
# round 2: host-pointer path with 3 / 7 / 15 row-copy helpers (config 5 and config 4)
mkdir -p gpurun_out
nproc
for t in 3 7 15; do
for wl in c5 c4; do
  HB_COPY_THREADS=$t timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-multi-hop --no-parity > gpurun_out/r2af_bench_${wl}_t$t.json 2> gpurun_out/r2af_bench_${wl}_t$t.err
done
done
python - <<'PY'
import json
for t in (3,7,15):
  for wl in ('c5','c4'):
    try:
        d=json.loads(open('gpurun_out/r2af_bench_%s_t%d.json'%(wl,t)).read().strip().splitlines()[-1])
        print('%s helpers %d value %.1f e2e %.1f' % (wl, t, d['value'], d['e2e']['value']))
    except Exception as e: print(wl,t,'failed',e)
PY
