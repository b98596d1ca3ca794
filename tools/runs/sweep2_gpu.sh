#!/bin/bash
TAG=${1:-rX}
source /dev/stdin <<'FN'
run() {
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
FN
mkdir -p gpurun_out
for st in 2 3 4 5; do run c4_serial_st$st HB_STAGES=$st -- --workload c4 --schedule serial; done
for st in 2 3 4; do run c4_over_st$st HB_STAGES=$st -- --workload c4; done
run c4_ldg1 X=1 -- --workload c4 --schedule serial --variant 0 --ctas-per-sm 1
run c4_ldg2 X=1 -- --workload c4 --schedule serial --variant 0 --ctas-per-sm 2
run c4_ldg4 X=1 -- --workload c4 --schedule serial --variant 0 --ctas-per-sm 4
for st in 2 3 4; do run c5_over_st$st HB_STAGES=$st -- --workload c5; done
run c5_serial_st3 HB_STAGES=3 -- --workload c5 --schedule serial
run c3_over X=1 -- --workload c3
