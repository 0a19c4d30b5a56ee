#!/usr/bin/env python
"""A reviewable artefact of a generated kernel (its .cu is a build product, not committed): ptxas resource usage, static
SASS instruction mix per function and an excerpt of the factorisation, extracted from the built library.
usage: kernel_artifact.py [kind] > profiles/<tag>_generated_<kind>_kernel.md"""
import collections, os, re, subprocess, sys, tempfile

kind = sys.argv[1] if len(sys.argv) > 1 else "trot"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "apf_quadruped_b200/lib/libqpswift_b200.so")], cwd=tmp, capture_output=True)
txt = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, f"warp_gen_{kind}.sm_100a.cubin")], capture_output=True, text=True).stdout.splitlines()
labels = [(i, l.strip().rstrip(":")) for i, l in enumerate(txt) if re.match(r"^\s*[\.\$_A-Za-z][^\s]*:\s*$", l) and (f"genw_{kind}" in l or "__internal" in l)]


def short(n):
    if "qp_solve_warp_kernel" in n and not n.startswith("$"):
        return "kernel body " + ("(CTA-synchronised)" if "ELb1E" in n else "(free-running)")
    m = re.search(rf"genw_{kind}\d+([a-z]+_\d+)E", n)
    if m:
        return m.group(1)
    m = re.search(r"__internal_\d+_\$(\w+)", n)
    return m.group(1) if m else n[:60]


out = [f"# Generated warp-per-QP kernel for the {kind} pattern", "",
       f"`build/gen/warp_gen_{kind}.cu` is a build product of `apf_quadruped_b200/codegen_warp.py` (not committed). Below: ptxas resource",
       "usage, the static SASS instruction mix per function and an excerpt of the factorisation (column steps of `LDL_numeric`:",
       "broadcast `LDS` of `L(j,i)`, `DMUL` + `DADD` per row -- unfused arithmetic is the bit-exactness contract -- `DFMA` only inside",
       "the division), extracted from `apf_quadruped_b200/lib/libqpswift_b200.so` with `cuobjdump -xelf` / `nvdisasm` (`scripts/kernel_artifact.py`).", "",
       f"## ptxas -v (`build/obj/gen_warp_gen_{kind}.ptxas.log`)", "", "```"]
out += [l.rstrip() for l in open(os.path.join(root, f"build/obj/gen_warp_gen_{kind}.ptxas.log")) if "Compiling" in l or "Used" in l or "spill" in l][:16]
out += ["```", "", "## static instruction mix per function (CTA-synchronised instantiation)", "",
        "| function | instructions | DMUL | DADD | DFMA | LDS | STS | LDG | SHFL | IMAD | MUFU | other |", "|---|---|---|---|---|---|---|---|---|---|---|---|"]
secs = [(name, i, labels[k + 1][0] if k + 1 < len(labels) else len(txt)) for k, (i, name) in enumerate(labels)]
for name, i, j in secs:
    if "ELb1E" not in name and "__internal" not in name:
        continue
    ops = collections.Counter()
    for l in txt[i:j]:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            ops[m.group(2)] += 1
    tot = sum(ops.values())
    if tot < 20:
        continue
    row = [short(name), tot] + [ops[k] for k in ("DMUL", "DADD", "DFMA", "LDS", "STS", "LDG", "SHFL", "IMAD", "MUFU")]
    row.append(tot - sum(row[2:]))
    out.append("| " + " | ".join(str(x) for x in row) + " |")
for name, i, j in secs:
    if "factor_1" in name and "ELb1E" in name:
        body = [l for l in txt[i:j] if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
        start = next(k for k, l in enumerate(body) if "DMUL" in l)
        out += ["", "## excerpt: `factor_1` (the second pass of rows: the dense tail), 70 instructions from the first DMUL", "", "```"]
        out += [re.sub(r"\s+", " ", l).strip()[:120] for l in body[start:start + 70]]
        out.append("```")
        break
print("\n".join(out))
