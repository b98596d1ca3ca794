#!/bin/bash
# round 2, final 8-GPU data: multi-GPU tests at 2 / 4 / 8 GPUs, then the bench at N = 8, 4, 2 (ours) and the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc; free -g | head -2
timeout 1200 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_multi.py tests/test_gpu_cpp_dropin.py -m gpu -x -q --durations=5 2>&1 | tail -30 > gpurun_out/r2s_pytest_multi.log
tail -12 gpurun_out/r2s_pytest_multi.log
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2s_bench_n$n.json 2> gpurun_out/r2s_bench_n$n.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2s_bench_n$n.json').read().strip().splitlines()[-1])
    print('N=$n value %.1f e2e %.1f (per-rank %.1f) ms/block %.4f kernel_ms %.4f hop_frac %.3f multi %s parity %s %s ts %s clk %s' % (d['value'], d['e2e']['value'], d['e2e'].get('per_rank_processes',{}).get('value',0), d['timing']['ms_per_block'], d['roofline']['kernel_ms'], d['roofline']['hop_frac'], [(m['blocks_per_call'], round(m['value'],1)) for m in d['multi_hop_reuse']['runs']], d['parity']['rel_rms'], d['parity'].get('rel_rms_host_pointer_path'), d['engine']['tail_streams'], d['clocks']['sm_mhz']))
except Exception as e:
    print('N=$n failed', e)
PY
  tail -3 gpurun_out/r2s_bench_n$n.err | cut -c1-300
done
python bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2s_bench_ref.json 2> gpurun_out/r2s_bench_ref.err; cut -c1-400 gpurun_out/r2s_bench_ref.json
