#!/bin/bash
# round 2, GPU call 2: packed multi-hop kernel -- parity, bench, ncu
mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv.py tests/test_gpu_fullsize.py -x -q -k "multi_hop or config4 or large_matrix or fft_size_change or skips_the_block or reset_of_one" 2>&1 | tail -15 > gpurun_out/r2g_pytest.log
cat gpurun_out/r2g_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2g_bench_c4.json 2> gpurun_out/r2g_bench_c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_c4.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['multi_hop_reuse'],d['parity']['rel_rms'], d['clocks'])
PY
tail -3 gpurun_out/r2g_bench_c4.err
for nh in 4 8; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cmac_mh2 -s 2 -c 1 -o gpurun_out/r2g_mh$nh -f \
    python tools/mh_probe.py $nh > gpurun_out/r2g_mh$nh.log 2>&1
ncu -i gpurun_out/r2g_mh$nh.ncu-rep --page details --csv > gpurun_out/r2g_mh${nh}_details.csv 2>/dev/null
ncu -i gpurun_out/r2g_mh$nh.ncu-rep --page raw --csv > gpurun_out/r2g_mh${nh}_raw.csv 2>/dev/null
tail -2 gpurun_out/r2g_mh$nh.log
done
