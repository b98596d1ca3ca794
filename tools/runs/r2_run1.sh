#!/bin/bash
# round 2, GPU call 1: the GPU suite with the full-size parity tests, then the bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err
tail -c 3000 gpurun_out/r2a_bench_c4.json; tail -5 gpurun_out/r2a_bench_c4.err
for wl in c1 c2 c3 c5; do
  python bench.py --workload $wl --steps 20 --warmup 5 --cpu-seconds 4 > gpurun_out/r2a_bench_$wl.json 2> gpurun_out/r2a_bench_$wl.err
  tail -c 1500 gpurun_out/r2a_bench_$wl.json; tail -3 gpurun_out/r2a_bench_$wl.err
done
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
cat gpurun_out/r2a_bench_ref.json; tail -3 gpurun_out/r2a_bench_ref.err
free -g | head -2; nproc
