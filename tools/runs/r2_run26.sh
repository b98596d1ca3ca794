#!/bin/bash
# round 2: final state again after the host-path changes (packed download, upload by copy kernel) -- whole suite, smoke(), config 5 / 4 lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2ar_pytest.log 2>&1; tail -3 gpurun_out/r2ar_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2ar_bench_c4.json 2> gpurun_out/r2ar_bench_c4.err
timeout 600 python bench.py --workload c5 --steps 20 --warmup 5 --cpu-seconds 4 > gpurun_out/r2ar_bench_c5.json 2> gpurun_out/r2ar_bench_c5.err
python - <<'PY'
import json
for wl in ('c4','c5'):
    try:
        d=json.loads(open('gpurun_out/r2ar_bench_%s.json'%wl).read().strip().splitlines()[-1])
        print('%s value %.1f e2e %.1f us/block %.2f hop_frac %.3f frac %.3f multi %s parity %.2e/%s cpu %.1f clk %s' % (wl, d['value'], d['e2e']['value'], d['timing']['ms_per_block']*1e3, d['roofline']['hop_frac'], d['roofline']['frac'], [(m['blocks_per_call'], round(m['value'],1)) for m in d['multi_hop_reuse']['runs']], d['parity']['rel_rms'], d['parity']['ok'], d['cpu_baseline']['value'], d['clocks']['sm_mhz']))
    except Exception as e: print(wl,'failed',e)
PY
