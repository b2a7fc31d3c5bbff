"""Per-source-line instruction counts from `ncu --page source --csv --print-source cuda,sass`.
Usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_lines.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        i_inst = [i for i, h in enumerate(r) if h == "Instructions Executed"][0]
        i_samp = [i for i, h in enumerate(r) if h == "# Samples"][0]
    elif hdr and r[0].isdigit():
        try:
            out.append((float(r[i_inst]), float(r[i_samp]), cur_file, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
tot = sum(o[0] for o in out)
tots = sum(o[1] for o in out)
print(f"total warp instructions {tot:.4g}, samples {tots:.0f}")
for inst, samp, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100 * inst / tot:5.1f}% inst {100 * samp / max(tots, 1):5.1f}% samp  {f}:{ln}  {src}")
