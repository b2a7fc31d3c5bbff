"""Summarise `ncu --page source --csv` output: stall reasons and SASS opcode mix.
Usage: ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_source_summary.py src.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) > ix["# Samples"]]


def f(r, k):
    try:
        return float(r[ix[k]] or 0)
    except (ValueError, IndexError):
        return 0.0


tot_inst = sum(f(r, "Instructions Executed") for r in data)
tot_samp = sum(f(r, "# Samples") for r in data)
print("SASS lines", len(data), "warp instructions", int(tot_inst), "samples", int(tot_samp))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {s:26s} {v:10.0f} {100 * v / max(tot_samp, 1):5.1f}%")
ops = collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    if m:
        ops[m.group(2).split(".")[0]] += f(r, "Instructions Executed")
for o, v in ops.most_common(22):
    print(f"  {o:12s} {v:14.0f} {100 * v / max(tot_inst, 1):5.1f}%")
