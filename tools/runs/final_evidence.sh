#!/bin/bash
# Run on the GPU box (under gpurun): the round's final captures and bench lines.  usage: tools/runs/final_evidence.sh <tag>
TAG=${1:-r1f}
mkdir -p gpurun_out
bash tools/profile_gpu.sh $TAG c5 c4
# the cluster transforms of config 5, full set, one launch each after warm-up
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_fwd_cl|k_inv_cl" -s 12 -c 2 -o gpurun_out/${TAG}_fftcl_c5 -f \
    python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu --no-multi-hop > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_fftcl_c5.ncu-rep --page details --csv > gpurun_out/${TAG}_fftcl_c5_details.csv
timeout 400 python bench.py > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
timeout 200 python bench.py --workload c5 > gpurun_out/${TAG}_bench_c5.json 2>/dev/null
timeout 300 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()"
for f in c4 c5 ref; do cut -c1-200 gpurun_out/${TAG}_bench_$f.json; done
