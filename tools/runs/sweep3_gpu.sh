#!/bin/bash
TAG=${1:-rX}
mkdir -p gpurun_out
run() {
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$name.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("%-22s value %8.2f  ms/step %.4f  e2e %8.2f  dom %.4f ms (%.3f of peak)  fwd %.4f head %.4f inv(+wait) %.4f  hop_frac %.3f  %s" % ("$name", d["value"], d["ms_per_step"], d["e2e"]["value"], r["kernel_ms"], r["frac"], r["forward_fft_ms"], r["head_cmac_ms"], r.get("inverse_fft_ms", r.get("wait_for_tail_plus_inverse_fft_ms")), r["hop_frac"], d["config"]["schedule"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("gpurun_out/${TAG}_$name.err").read()[-800:])
PY
}
for rep in 1 2 3; do
run c4_auto_$rep X=1 -- --workload c4
run c4_serial_$rep X=1 -- --workload c4 --schedule serial
run c4_over_st4_$rep HB_STAGES=4 -- --workload c4
run c4_over_rs16_$rep HB_RESERVE=16 -- --workload c4
done
for w in c1 c2 c3 c5; do run ${w}_auto X=1 -- --workload $w; done
run c5_serial X=1 -- --workload c5 --schedule serial
run c4_hops4 X=1 -- --workload c4 --hops 4
