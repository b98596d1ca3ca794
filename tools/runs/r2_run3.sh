#!/bin/bash
# round 2, 2-GPU call: the whole GPU suite (multi-device front, sharded engine at 2 ranks, full-size parity), then N=2 bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2h_pytest.log
tail -25 gpurun_out/r2h_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
tail -c 2500 gpurun_out/r2h_bench_n2.json; tail -5 gpurun_out/r2h_bench_n2.err
