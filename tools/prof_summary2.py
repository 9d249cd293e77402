#!/usr/bin/env python3
"""Summarise an ncu report (source page + raw page) by kernel phase and by source line.  Runs on the CPU box:
    python tools/prof_summary2.py gpurun_out/prof_x.ncu-rep <kernel source .cuh> [out.md]
Phases = the `// ---- X:` comment markers of the kernel source."""
import collections, csv, io, re, subprocess, sys

rep, ksrc = sys.argv[1], sys.argv[2]
out = open(sys.argv[3], "w") if len(sys.argv) > 3 else sys.stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
rawrows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rawrows[0], rawrows[-1]
m = dict(zip(hdr, vals))
keys = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]
print("| metric | value |\n|---|---|", file=out)
for k in keys:
    if k in m:
        print(f"| {k} | {m[k]} |", file=out)
for k in sorted(m):
    if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and k not in keys:
        try:
            if float(m[k]) >= 0.05:
                print(f"| {k} | {m[k]} |", file=out)
        except ValueError:
            pass
kfile = ksrc.split("/")[-1]
kernel_src = open(ksrc).read().split("\n")
marks = [("helpers / setup", 1)]
for i, l in enumerate(kernel_src):
    mm = re.match(r"\s*// ---- (.*?)[-\s]*$", l)
    if mm:
        marks.append((mm.group(1)[:70], i + 1))
    elif "// Persistent" in l:
        marks.append(("persistent loop", i + 1))
    elif "// pass 1:" in l or "// pass 2:" in l or "// pass 3:" in l or "// splitter list ranking" in l or "// thread the runs" in l:
        marks.append((l.strip()[3:70], i + 1))


def phase(line):
    cur = "?"
    for name, start in marks:
        if line >= start:
            cur = name
    return cur


data, agg, fn = [], collections.OrderedDict(), None
for r in rows:
    if r and r[0] == "File Path":
        fn = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0].isdigit() and r[2] == "-":
        try:
            s, i = int(r[4] or 0), int(r[7] or 0)
        except ValueError:
            continue
        data.append((s, i, fn, int(r[0]), r[1].strip()[:110]))
        k = phase(int(r[0])) if fn == kfile else fn
        a = agg.setdefault(k, [0, 0]); a[0] += s; a[1] += i
tot = sum(d[0] for d in data) or 1
toti = sum(d[1] for d in data) or 1
print(f"\nstall samples {tot}, warp instructions {toti}\n\n| phase | samples | instructions |\n|---|---|---|", file=out)
for k, (s, i) in agg.items():
    print("| %s | %.1f%% | %.1f%% |" % (k, 100 * s / tot, 100 * i / toti), file=out)
print("\n```", file=out)
for d in sorted(data, reverse=True)[:40]:
    print("%5.1f%% smp %5.1f%% inst %s:%-4d %s" % (100 * d[0] / tot, 100 * d[1] / toti, d[2][:16], d[3], d[4]), file=out)
print("```", file=out)
