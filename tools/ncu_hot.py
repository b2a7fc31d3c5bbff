"""Hot source lines / SASS opcode mix of one kernel from an .ncu-rep (source page, cuda+sass view).
Usage: python tools/ncu_hot.py rep.ncu-rep 'kernel name substring' [top]"""
import collections
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
# function names on the source page look like `void tqb_pass_kernel<(int)4, (bool)1, (bool)1>(PassArgs)`
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
allrows = list(csv.reader(out.splitlines()))
# keep the first kernel section whose function name contains `kre` (plain substring)
rows, keep, seen = [], False, False
for i, r in enumerate(allrows):
    if r and r[0] == "Function Name":
        match = kre in r[1]
        if match and seen and not keep:
            break
        if keep and not match:
            break
        if match and not seen:
            seen = keep = True
            rows.append(allrows[i - 1])
    if keep:
        rows.append(r)
cur_file, cur_line, cur_src, hdr = None, None, "", None
lines = collections.defaultdict(lambda: [0.0, 0.0, ""])
opc = collections.defaultdict(float)
stalls = collections.defaultdict(float)
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        i_inst = [i for i, h in enumerate(r) if h == "Instructions Executed"][0]
        i_samp = [i for i, h in enumerate(r) if h == "# Samples"][0]
        stall_cols = [(i, h) for i, h in enumerate(r) if h.startswith("stall_") and "Not Issued" not in h]
    elif hdr and r[0].isdigit():
        cur_line, cur_src = int(r[0]), r[1].strip()[:100]
    elif hdr and r[0] == "" and len(r) > i_inst and r[2].startswith("0x"):
        try:
            inst, samp = float(r[i_inst]), float(r[i_samp])
        except ValueError:
            continue
        k = (cur_file, cur_line)
        lines[k][0] += inst
        lines[k][1] += samp
        lines[k][2] = cur_src
        op = r[3].split()[0] if not r[3].strip().startswith("@") else r[3].split()[1]
        opc[op.split(".")[0]] += inst
        for i, h in stall_cols:
            try:
                stalls[h] += float(r[i])
            except ValueError:
                pass
tot = sum(v[0] for v in lines.values())
tots = sum(v[1] for v in lines.values())
print(f"total warp instructions {tot:.4g}, samples {tots:.0f}")
print("opcode mix:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(opc.items(), key=lambda kv: -kv[1])[:18]))
ts = sum(stalls.values())
print("stall mix:", ", ".join(f"{k[6:]} {100 * v / ts:.1f}%" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]))
for (f, ln), (inst, samp, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * inst / tot:5.1f}% inst {100 * samp / max(tots, 1):5.1f}% samp  {f}:{ln}  {src}")
