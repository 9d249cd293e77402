#!/usr/bin/env python3
"""Static SASS instruction count per source line of one kernel (no GPU needed).
    python tools/sass_lines.py <lib.so> <kernel-name-substring> [file-substring] [first-line last-line]
Uses cuobjdump -xelf + nvdisasm -g (line info from -lineinfo)."""
import collections, os, re, subprocess, sys, tempfile
lib, pat = sys.argv[1], sys.argv[2]
fsub = sys.argv[3] if len(sys.argv) > 3 else ""
lo, hi = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 1 << 30)
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin") and "ingest" not in f][0]
out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
cnt, tot, infn, cur = collections.Counter(), 0, False, None
for l in out.split("\n"):
    m = re.match(r"\s*\.text\.(\S+):", l) or re.match(r"\s*//-+ \.text\.(\S+)", l)
    if m:
        infn = pat in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+\S", l):
        tot += 1
        if cur:
            cnt[cur] += 1
print("total SASS instructions:", tot)
sel = sorted((k, v) for k, v in cnt.items() if fsub in k[0] and lo <= k[1] <= hi)
print("selected range:", sum(v for _, v in sel))
for (f, ln), v in sel:
    print(f"{f}:{ln}\t{v}")
