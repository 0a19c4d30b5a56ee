import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, lib, kname = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = []
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        infn, where = False, ("?", 0, "")
        for ln in out.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                infn = kname in m.group(1); continue
            if not infn: continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
                where = (os.path.basename(m.group(1)), int(m.group(2)), " <- ".join(f"{os.path.basename(a)}:{b}" for a,b in inl)); continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m: lines.append((where, m.group(2).strip()))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [(int(r[isamp]), int(r[iex])) for r in rows[2:] if len(r) > iex]
assert len(data) == len(lines)
agg = defaultdict(lambda: [0,0,0])
for ((f,n,inl),sass),(smp,ex) in zip(lines,data):
    if f.startswith("warp_gen"): key=("gen",0,"")
    else: key=(f,n,inl)
    a=agg[key]; a[0]+=smp; a[1]+=ex; a[2]+=1
ts=sum(a[0] for a in agg.values()); te=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:70]:
    print(f"{100*a[0]/ts:5.2f}% smp {100*a[1]/te:5.2f}% ex {a[2]:4d} static  {k[0]}:{k[1]} {k[2][:90]}")
