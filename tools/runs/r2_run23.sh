#!/bin/bash
# round 2 (2 GPUs): bench at N = 2 after the calibration decision became collective
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2ak_bench_n2.json 2> gpurun_out/r2ak_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2ak_bench_n2.json').read().strip().splitlines()[-1])
    print('N=2 value %.1f e2e %.1f ms/block %.4f parity %s multi %s cpu %s' % (d['value'], d['e2e']['value'], d['timing']['ms_per_block'], d['parity'].get('rel_rms'), [(m['blocks_per_call'], round(m['value'],1)) for m in d['multi_hop_reuse']['runs']], d['cpu_baseline']['value']))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2ak_bench_n2.err').read()[-1500:])
PY
