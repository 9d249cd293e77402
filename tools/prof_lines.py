#!/usr/bin/env python3
"""Per-source-line dynamic warp-instruction counts (and stall samples) of an ncu report, divided by a unit count:
    python tools/prof_lines.py rep.ncu-rep <file-substr> <units> [min_inst_per_unit]"""
import csv, io, subprocess, sys
rep, fsub, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
fn, tot = None, 0
rows = []
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "File Path":
        fn = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0].isdigit() and r[2] == "-":
        try: s, i = int(r[4] or 0), int(r[7] or 0)
        except ValueError: continue
        tot += i
        if fsub in (fn or ""): rows.append((int(r[0]), i / units, s, r[1].strip()[:120]))
print("total inst/unit: %.1f" % (tot / units))
for ln, i, s, t in rows:
    if i >= thr: print("%5d %8.1f %7d  %s" % (ln, i, s, t))
