#!/bin/bash
# Run on the GPU box (under gpurun): launch lists and one full ncu capture of the dominant kernel.
# usage: tools/profile_gpu.sh <tag> [workloads...]      outputs under gpurun_out/
TAG=${1:-rX}; shift
WLS=${@:-c4}
mkdir -p gpurun_out
KREGEX='regex:^k_(fwd|cmac|inv|rows|gather|td|hop)'
for WL in $WLS; do
  # every launch of the hop kernels with its device time (shares, not absolutes)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 60 --csv \
      --log-file gpurun_out/${TAG}_launches_${WL}.csv python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu \
      > gpurun_out/${TAG}_launches_${WL}.log 2>&1
  # the multiply-accumulate kernel, full set, one launch after warm-up
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cmac -s 4 -c 1 \
      -o gpurun_out/${TAG}_cmac_${WL} -f python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu \
      > gpurun_out/${TAG}_cmac_${WL}.log 2>&1
  ncu -i gpurun_out/${TAG}_cmac_${WL}.ncu-rep --page raw --csv > gpurun_out/${TAG}_cmac_${WL}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_cmac_${WL}.ncu-rep --page details --csv > gpurun_out/${TAG}_cmac_${WL}_details.csv 2>/dev/null
done
