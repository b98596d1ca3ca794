#!/usr/bin/env python
"""Condense ncu exports (made on the GPU box by tools/profile_gpu.sh) into the small text files kept under profiles/.

    python tools/ncu_summary.py details <details.csv> <out.txt> "<title>"
    python tools/ncu_summary.py traffic <raw.csv> <key>            # updates profiles/traffic.json
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def details(src, dst, title):
    rows = list(csv.reader(open(src)))
    hdr = rows[0]
    ix = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Section Name", "Metric Name", "Metric Unit", "Metric Value")}
    out = ["# " + title]
    seen = set()
    for r in rows[1:]:
        if len(r) <= ix["Metric Value"] or not r[ix["Metric Name"]]:
            continue
        key = (r[ix["ID"]], r[ix["Section Name"]], r[ix["Metric Name"]])
        if key in seen:
            continue
        seen.add(key)
        if ("kernel", r[ix["ID"]]) not in seen:
            seen.add(("kernel", r[ix["ID"]]))
            out.append("## launch %s: %s" % (r[ix["ID"]], r[ix["Kernel Name"]]))
        out.append("%s | %s | %s | %s" % (r[ix["Section Name"]], r[ix["Metric Name"]], r[ix["Metric Unit"]], r[ix["Metric Value"]]))
    open(dst, "w").write("\n".join(out) + "\n")


SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def traffic(src, key):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    tot = []
    for vals in rows[2:]:
        b = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            k = hdr.index(name)
            b += float(vals[k].replace(",", "")) * SCALE[units[k]]
        tot.append(b)
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(path))
    except Exception:
        d = {}
    d[key] = sum(tot) / len(tot)
    # provenance: which capture the figure came from and when it was condensed (bench.py prints it as roofline.traffic_source)
    import datetime
    d.setdefault("_source", {})[key] = "%s (ncu --set full, %d launch(es); condensed %s)" % (os.path.basename(src), len(tot), datetime.date.today().isoformat())
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)
    print(key, d[key])


if __name__ == "__main__":
    if sys.argv[1] == "details":
        details(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        traffic(sys.argv[2], sys.argv[3])
