#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2r_bench_c4.json 2> gpurun_out/r2r_bench_c4.err
python bench.py --workload c5 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2r_bench_c5.json 2> gpurun_out/r2r_bench_c5.err
python bench.py --workload c4r8 --steps 20 --warmup 5 --no-cpu --no-parity > gpurun_out/r2r_bench_c4r8.json 2> gpurun_out/r2r_bench_c4r8.err
python - <<'PY'
import json
for wl in ('c4','c5','c4r8'):
    d=json.loads(open('gpurun_out/r2r_bench_%s.json'%wl).read().strip().splitlines()[-1])
    print('%s value %.1f e2e %.1f us/block %.2f hop_frac %.3f kernel_ms %.4f frac %.3f multi %s clk %s' % (wl, d['value'], d['e2e']['value'], d['timing']['ms_per_block']*1e3, d['roofline']['hop_frac'], d['roofline']['kernel_ms'], d['roofline']['frac'], [(m['blocks_per_call'], round(m['value'],1)) for m in d['multi_hop_reuse']['runs']], d['clocks']['sm_mhz']))
PY
timeout 600 ncu --set full --clock-control none -k regex:k_cmac_tma -s 6 -c 1 -o gpurun_out/r2r_cmac_c4 -f python bench.py --workload c4 --steps 2 --warmup 3 --blocks-per-step 4 --no-cpu --no-parity --no-multi-hop --tail-streams 1 > gpurun_out/r2r_cmac_c4.log 2>&1
ncu -i gpurun_out/r2r_cmac_c4.ncu-rep --page details --csv > gpurun_out/r2r_cmac_c4_details.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2r_cmac_c4_details.csv')))
hdr=rows[0]; i_m=hdr.index('Metric Name'); i_v=hdr.index('Metric Value')
print(' '.join('%s=%s'%(r[i_m],r[i_v]) for r in rows[1:] if r[i_m] in ('Duration','DRAM Throughput','Registers Per Thread','Issue Slots Busy')))
PY
