#!/bin/bash
# round 2 (2 GPUs): multi-device tests incl. the fused exchange at four-step FFT sizes, bench sanity at N = 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -m gpu -x -q --durations=5 > gpurun_out/r2as_pytest_multi.log 2>&1; tail -12 gpurun_out/r2as_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2as_bench_n2.json 2> gpurun_out/r2as_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2as_bench_n2.json').read().strip().splitlines()[-1])
    print('N=2 value %.1f e2e %.1f ms/block %.4f parity %s multi %s' % (d['value'], d['e2e']['value'], d['timing']['ms_per_block'], d['parity'].get('rel_rms'), [(m['blocks_per_call'], round(m['value'],1)) for m in d['multi_hop_reuse']['runs']]))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2as_bench_n2.err').read()[-1500:])
PY
