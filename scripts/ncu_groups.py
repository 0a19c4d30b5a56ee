#!/usr/bin/env python
"""Attribute ncu's per-SASS-instruction samples to source files / generated functions / line ranges.
usage: ncu_groups.py <prof.ncu-rep> <lib.so> <kernel-substring> [generated.cu]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, lib, kname = sys.argv[1:4]
gen = sys.argv[4] if len(sys.argv) > 4 else None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = []
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        infn, where = False, ("?", 0)
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
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [(int(r[isamp]), int(r[iex])) for r in rows[2:] if len(r) > iex]
assert len(data) == len(lines), (len(data), len(lines))
fn = {}
if gen:
    cur = None
    for i, l in enumerate(open(gen).read().splitlines(), 1):
        m = re.match(r"static __device__ __(?:noinline|forceinline)__ \w+ (\w+)\(", l)
        if m:
            cur = m.group(1)
        fn[i] = cur
RANGES = {"kernels.cuh": [(283, 326, "ipm_residuals"), (328, 353, "steplength"), (355, 363, "ipm_rhs"), (382, 423, "init_point"), (425, 434, "lambda_mu"),
                          (436, 440, "update_kkt"), (442, 452, "recover_dz_ds"), (454, 466, "ipm_sigma"), (99, 127, "copy_in_out"), (136, 140, "regularise")]}
agg = defaultdict(lambda: [0, 0, 0])
ops = defaultdict(lambda: [0, 0])
for ((f, n), sass), (smp, ex) in zip(lines, data):
    key = f
    if gen and f == os.path.basename(gen):
        key = "gen:" + str(fn.get(n))
    for lo, hi, name in RANGES.get(f, []):
        if lo <= n <= hi:
            key = f + ":" + name
    a = agg[key]
    a[0] += smp; a[1] += ex; a[2] += 1
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    op = op.split(".")[0]
    ops[op][0] += smp; ops[op][1] += ex
ts = sum(a[0] for a in agg.values()); te = sum(a[1] for a in agg.values())
print(f"total samples {ts}, warp instructions executed {te}, static instructions {len(data)}")
print("| group | samples | executed | static |\n|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| {k} | {100 * a[0] / ts:.1f}% | {100 * a[1] / te:.1f}% | {a[2]} |")
print("\n| opcode | samples | executed |\n|---|---|---|")
for k, a in sorted(ops.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"| {k} | {100 * a[0] / ts:.1f}% | {100 * a[1] / te:.1f}% |")
