#!/bin/bash
# round 2, 2-GPU call: the GPU suite again (sharded multi-hop batches, hop batches on small engines, automatic tail streams), benches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2j_pytest.log
tail -12 gpurun_out/r2j_pytest.log
for wl in c1 c2 c3; do
  python bench.py --workload $wl --steps 20 --warmup 5 --cpu-seconds 3 > gpurun_out/r2j_bench_$wl.json 2> gpurun_out/r2j_bench_$wl.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2j_bench_$wl.json').read().strip().splitlines()[-1])
print('$wl value %.1f e2e %.1f ms/block %.4f multi %s parity %s cpu %.1f' % (d['value'], d['e2e']['value'], d['timing']['ms_per_block'], [(m['blocks_per_call'], round(m['value'],1)) for m in d['multi_hop_reuse']['runs']], d['parity']['ok'], d['cpu_baseline']['value']))
PY
  tail -2 gpurun_out/r2j_bench_$wl.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2j_bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value %.1f e2e %.1f ms/block %.4f multi %s parity %s' % (d['value'], d['e2e']['value'], d['timing']['ms_per_block'], [(m['blocks_per_call'], round(m['value'],1)) for m in d['multi_hop_reuse']['runs']], d['parity']))
PY
tail -3 gpurun_out/r2j_bench_n2.err
