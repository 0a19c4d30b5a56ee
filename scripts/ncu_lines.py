#!/usr/bin/env python
"""Attribute ncu's per-SASS-instruction samples to CUDA source lines.
usage: ncu_lines.py <prof.ncu-rep> <lib.so> <kernel-substring> [top]
Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` line info of the same binary."""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = []
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, infn, where = None, False, ("?", 0)
        for ln in out.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                infn = kname in m.group(1)
                continue
            if not infn:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                where = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                lines.append((where, m.group(2).strip()))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [(r[isrc].strip(), int(r[isamp]), int(r[iex])) for r in rows[2:] if len(r) > iex]
if len(data) != len(lines):
    print(f"warning: {len(data)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
agg = defaultdict(lambda: [0, 0, 0])
for (where, _), (_, smp, ex) in zip(lines, data):
    a = agg[where]
    a[0] += smp; a[1] += ex; a[2] += 1
ts = sum(a[0] for a in agg.values()); te = sum(a[1] for a in agg.values())
print(f"total samples {ts}, warp instructions executed {te}, static instructions {len(data)}")
src_cache = {}
def text(where):
    f, n = where
    for root in ("apf_quadruped_b200/csrc", "."):
        pth = os.path.join(root, f)
        if os.path.exists(pth):
            src_cache.setdefault(pth, open(pth).read().splitlines())
            L = src_cache[pth]
            return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
    return ""
print("| file:line | samples | executed | static | source |\n|---|---|---|---|---|")
for where, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"| {where[0]}:{where[1]} | {100 * a[0] / ts:.1f}% | {100 * a[1] / te:.1f}% | {a[2]} | `{text(where)}` |")
