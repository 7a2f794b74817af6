#!/usr/bin/env python
"""Record the DRAM traffic of the profiled kernels (one `ncu --set full` capture, raw page CSV)
in profiles/traffic.json, from where bench.py fills `roofline.traffic`.
Usage: ncu_traffic.py <raw.csv> <workload> "<pass label>=<row id>" ...
(row id = the ID column of the raw page; pass labels are bench.py's pass names)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, workload, maps = sys.argv[1], sys.argv[2], sys.argv[3:]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(v.replace(",", "")) * scale


path = os.path.join(ROOT, "profiles", "traffic.json")
db = json.load(open(path)) if os.path.exists(path) else {}
entry = db.setdefault(workload, {})
for m in maps:
    label, rid = m.rsplit("=", 1)
    r = next(r for r in rows[2:] if r[col["ID"]] == rid)
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    entry[label] = {"dram_bytes": int(rd + wr), "read": int(rd), "write": int(wr),
                    "kernel": r[col["Kernel Name"]], "capture": os.path.relpath(raw, ROOT)}
json.dump(db, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(entry, indent=1))
