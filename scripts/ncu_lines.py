"""Per-source-line share of executed warp instructions and stall samples from an ncu capture taken with
--import-source on (kernels compiled with -lineinfo).  usage: ncu_lines.py report.ncu-rep [min_pct]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fname, ie, isamp, agg = None, None, None, {}
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        ie, isamp = r.index("Instructions Executed"), r.index("# Samples")
        continue
    if r[0] in ("", "-") or r[2] != "-":
        continue
    try:
        n, s = int(r[ie]), int(r[isamp])
    except ValueError:
        continue
    k = (fname, int(r[0]))
    a = agg.get(k, (0, 0, r[1]))
    agg[k] = (a[0] + n, a[1] + s, r[1])
tot = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print(f"{tot} warp instructions, {ts} stall samples")
for k, v in sorted(agg.items()):
    if v[0] > tot * min_pct / 100 or v[1] > ts * min_pct / 100:
        print(f"{k[0][:16]:16s} {k[1]:4d} {100 * v[0] / tot:5.1f}% instr {100 * v[1] / ts:5.1f}% samples  {v[2].strip()[:100]}")
