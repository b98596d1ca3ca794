#!/bin/bash
# round 2, final 1-GPU data: the bench lines of every config and the reference arm
mkdir -p gpurun_out
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2q_bench_ref.json 2> gpurun_out/r2q_bench_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench_c4.json 2> gpurun_out/r2q_bench_c4.err
for wl in c1 c2 c3 c5; do
  python bench.py --workload $wl --steps 20 --warmup 5 --cpu-seconds 4 > gpurun_out/r2q_bench_$wl.json 2> gpurun_out/r2q_bench_$wl.err
done
python - <<'PY'
import json
for wl in ('ref','c4','c1','c2','c3','c5'):
    try:
        d=json.loads(open('gpurun_out/r2q_bench_%s.json'%wl).read().strip().splitlines()[-1])
        if wl=='ref': print('ref value %.2f cores %s sample %s' % (d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample'][:60])); continue
        print('%s value %.1f e2e %.1f us/block %.2f hop_frac %.3f frac %.3f share %.3f multi %s parity %.2e/%s cpu %.1f clk %s ts %s' % (wl, d['value'], d['e2e']['value'], d['timing']['ms_per_block']*1e3, d['roofline']['hop_frac'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], [(m['blocks_per_call'], round(m['value'],1)) for m in d['multi_hop_reuse']['runs']], d['parity']['rel_rms'], d['parity']['ok'], d['cpu_baseline']['value'], d['clocks']['sm_mhz'], d['engine']['tail_streams']))
    except Exception as e: print(wl,'failed',e)
PY
