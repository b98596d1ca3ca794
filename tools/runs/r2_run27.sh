#!/bin/bash
# round 2 (4 GPUs): bench at N = 4, where 16 calibration blocks take about the 5 ms at which the calibration stops (the decision is collective)
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2at_bench_n4.json 2> gpurun_out/r2at_bench_n4.err
echo "exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2at_bench_n4.json').read().strip().splitlines()[-1])
    print('N=4 value %.1f e2e %.1f ms/block %.4f parity %s multi %s' % (d['value'], d['e2e']['value'], d['timing']['ms_per_block'], d['parity'].get('rel_rms'), [(m['blocks_per_call'], round(m['value'],1)) for m in d['multi_hop_reuse']['runs']]))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2at_bench_n4.err').read()[-1200:])
PY
