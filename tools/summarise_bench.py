"""one line per bench.py JSON file: python tools/summarise_bench.py file.json ..."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print("%-34s %-10s value %9.2f  ms/step %.4f  e2e %9.2f  frac %s  hop_frac %s  cpu %s  multi-hop %s" % (
            path.split("/")[-1], d["config"].get("schedule", d.get("impl", "")), d["value"], d["ms_per_step"], d["e2e"]["value"],
            "%.3f" % r["frac"] if "frac" in r else "-", "%.3f" % r["hop_frac"] if "hop_frac" in r else "-",
            "%.2f" % d["cpu_baseline"]["value"] if d.get("cpu_baseline") else "-",
            "%.1f" % d["multi_hop_reuse"]["value"] if d.get("multi_hop_reuse") else "-"))
    except Exception as e:
        print(path, "FAILED", e)
