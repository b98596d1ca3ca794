#!/bin/bash
# round 2: confirmation of HEAD on one GPU -- whole suite, synccheck of the multi-hop kernel's named barrier, bench lines of every config
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2ae_pytest.log 2>&1; tail -4 gpurun_out/r2ae_pytest.log
echo "### synccheck multi-hop kernel"
timeout 200 compute-sanitizer --tool synccheck --print-limit 3 --error-exitcode 9 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k 'multi_hop_reuse and float32-None' 2>&1 | grep -E "========= (Barrier|Error|ERROR)|     at |passed|failed" | head -6 | tee gpurun_out/r2ae_synccheck_mh.txt
nvcc -O2 -o /tmp/hop_rate tools/hop_rate.cu -Iinclude -Lhisstools_library_b200/lib -lhisstools_b200 -Xlinker -rpath=$PWD/hisstools_library_b200/lib -Wno-deprecated-gpu-targets \
  && timeout 300 /tmp/hop_rate > gpurun_out/r2ae_hop_rate.txt 2>&1
cat gpurun_out/r2ae_hop_rate.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2ae_bench_c4.json 2> gpurun_out/r2ae_bench_c4.err
for wl in c1 c2 c3 c5; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --cpu-seconds 4 > gpurun_out/r2ae_bench_$wl.json 2> gpurun_out/r2ae_bench_$wl.err
done
python - <<'PY'
import json
for wl in ('c4','c1','c2','c3','c5'):
    try:
        d=json.loads(open('gpurun_out/r2ae_bench_%s.json'%wl).read().strip().splitlines()[-1])
        print('%s value %.1f e2e %.1f us/block %.2f hop_frac %.3f frac %.3f share %.3f multi %s parity %.2e/%s cpu %.1f clk %s' % (wl, d['value'], d['e2e']['value'], d['timing']['ms_per_block']*1e3, d['roofline']['hop_frac'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], [(m['blocks_per_call'], round(m['value'],1)) for m in d['multi_hop_reuse']['runs']], d['parity']['rel_rms'], d['parity']['ok'], d['cpu_baseline']['value'], d['clocks']['sm_mhz']))
    except Exception as e: print(wl,'failed',e)
PY
