#!/bin/bash
# round 2, 1-GPU call: full suite, exchange probe, host-path check (config 5), ncu captures for profiles/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2l_pytest.log; tail -4 gpurun_out/r2l_pytest.log
./tools/fft_exchange_probe > gpurun_out/r2l_exchange_probe.txt 2>&1; cat gpurun_out/r2l_exchange_probe.txt
python bench.py --workload c5 --steps 20 --warmup 5 --cpu-seconds 3 > gpurun_out/r2l_bench_c5.json 2> gpurun_out/r2l_bench_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l_bench_c5.json').read().strip().splitlines()[-1])
print('c5 value %.1f e2e %.1f ms/block %.4f e2e ms/block %.4f parity %s' % (d['value'], d['e2e']['value'], d['timing']['ms_per_block'], d['e2e']['ms_per_block'], d['parity']['ok']))
PY
KREGEX='regex:^k_(fwd|cmac|inv|rows|gather|td|hop)'
for WL in c4 c4r8 c5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 80 --csv \
      --log-file gpurun_out/r2l_launches_${WL}.csv python bench.py --workload $WL --steps 2 --warmup 3 --blocks-per-step 4 --no-cpu --no-parity --no-multi-hop \
      > gpurun_out/r2l_launches_${WL}.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cmac_tma -s 6 -c 1 \
      -o gpurun_out/r2l_cmac_${WL} -f python bench.py --workload $WL --steps 2 --warmup 3 --blocks-per-step 4 --no-cpu --no-parity --no-multi-hop --tail-streams 1 \
      > gpurun_out/r2l_cmac_${WL}.log 2>&1
  ncu -i gpurun_out/r2l_cmac_${WL}.ncu-rep --page raw --csv > gpurun_out/r2l_cmac_${WL}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2l_cmac_${WL}.ncu-rep --page details --csv > gpurun_out/r2l_cmac_${WL}_details.csv 2>/dev/null
  tail -1 gpurun_out/r2l_cmac_${WL}.log | cut -c1-200
done
