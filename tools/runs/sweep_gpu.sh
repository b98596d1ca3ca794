#!/bin/bash
# usage: tools/runs/sweep_gpu.sh <tag>   -- schedule / ring-depth comparison on one GPU (bench lines under gpurun_out/)
TAG=${1:-rX}
mkdir -p gpurun_out
run() { # name, env..., -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$name.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("%-26s value %8.2f  ms/step %.4f  e2e %8.2f  dom %.4f ms (%.3f of peak)  fwd %.4f head %.4f inv(+wait) %.4f  hop_frac %.3f  %s" % ("$name", d["value"], d["ms_per_step"], d["e2e"]["value"], r["kernel_ms"], r["frac"], r["forward_fft_ms"], r["head_cmac_ms"], r.get("inverse_fft_ms", r.get("wait_for_tail_plus_inverse_fft_ms")), r["hop_frac"], d["config"]["schedule"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("gpurun_out/${TAG}_$name.err").read()[-800:])
PY
}
run c4_serial X=1 -- --workload c4 --schedule serial
run c4_serial_st4 HB_STAGES=4 -- --workload c4 --schedule serial
run c4_serial_st3 HB_STAGES=3 -- --workload c4 --schedule serial
run c4_over X=1 -- --workload c4
run c4_over_st3 HB_STAGES=3 -- --workload c4
run c4_over_rs16 HB_STAGES=6 HB_RESERVE=16 -- --workload c4
run c4_over_rs32 HB_STAGES=6 HB_RESERVE=32 -- --workload c4
run c5_serial X=1 -- --workload c5 --schedule serial
run c5_over X=1 -- --workload c5
run c5_over_rs32 HB_RESERVE=32 -- --workload c5
run c3_serial X=1 -- --workload c3 --schedule serial
run c3_over X=1 -- --workload c3
run c2_serial X=1 -- --workload c2 --schedule serial
run c2_over X=1 -- --workload c2
run c1_serial X=1 -- --workload c1 --schedule serial
run c1_over X=1 -- --workload c1
