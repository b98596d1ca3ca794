#!/bin/bash
# round 2: final state on one GPU -- whole suite, smoke(), call rates (device and host rows, C++ caller), default bench line, ncu of the fused hop
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2aj_pytest.log 2>&1; tail -3 gpurun_out/r2aj_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
nvcc -O2 -o /tmp/hop_rate tools/hop_rate.cu -Iinclude -Lhisstools_library_b200/lib -lhisstools_b200 -Xlinker -rpath=$PWD/hisstools_library_b200/lib -Wno-deprecated-gpu-targets \
  && timeout 300 /tmp/hop_rate > gpurun_out/r2aj_hop_rate.txt 2>&1
cat gpurun_out/r2aj_hop_rate.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2aj_bench_c4.json 2> gpurun_out/r2aj_bench_c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2aj_bench_c4.json').read().strip().splitlines()[-1])
print('c4 value %.1f e2e %.1f ms/block %.4f hop_frac %.3f frac %.3f parity %.2e cpu %.2f clk %s' % (d['value'], d['e2e']['value'], d['timing']['ms_per_block'], d['roofline']['hop_frac'], d['roofline']['frac'], d['parity']['rel_rms'], d['cpu_baseline']['value'], d['clocks']['sm_mhz']))
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hop_fused -s 20 -c 2 -o gpurun_out/r2aj_hop_fused_c3 -f \
    python bench.py --workload c3 --steps 3 --warmup 3 --blocks-per-step 16 --no-cpu --no-multi-hop --no-parity --hop-overlap 0 > gpurun_out/r2aj_ncu_c3.log 2>&1
ncu -i gpurun_out/r2aj_hop_fused_c3.ncu-rep --page details --csv > gpurun_out/r2aj_hop_fused_c3_details.csv 2>/dev/null
ls -la gpurun_out/r2aj_hop_fused_c3* | head
