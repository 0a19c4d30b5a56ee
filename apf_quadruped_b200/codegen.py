"""Pattern-specialised throughput kernels: the micro-programs of csrc/stream_program.cpp turned into
straight-line CUDA.

The interpreter of csrc/stream.cuh spends ~20 instructions per micro-op on instruction supply (shuffles,
flag tests, address arithmetic); for a KNOWN sparsity pattern every index and every flag is a constant,
so the same micro-op sequence can be emitted as source: one `ld.global` with an immediate offset, one
`ld.shared` with an immediate offset, one multiply and one subtract per term.  The arithmetic (and its
order) is the program's, word for word -- the programs themselves are checked against the reference on
the host (tests/test_stream_program.py) -- so the generated kernels stay bit-identical with reference
qpSWIFT (ldl.c:253-326 LDL_numeric, 495-597 the solves, Auxilary.c:745-786 computeresiduals).

    python -m apf_quadruped_b200.codegen --out build/gen trot stance crawl

writes build/gen/stream_gen_<kind>.cu; the Makefile compiles whatever it finds there into the library,
where each file registers its kernel under the key of its micro-program (QPB_PLAN_STREAM_KEY).  A
pattern without a generated kernel runs on the interpreter.
"""
from __future__ import annotations

import argparse
import os

TERM, DIVACC, DIV, INITD, ENDD, LOADS, NOP = 0, 2, 3, 4, 5, 6, 8
FIRSTZ, INIT, ST_SHARED, ST_HBM, LATE1, ONE, SLOW = (1 << (52 + i) for i in range(7))


class Op:
    __slots__ = ("f1", "f2", "c", "op", "w", "uses_p", "uses_s", "load_at", "name")

    def __init__(self, w: int):
        self.w = w
        self.f1 = w & 0xffff
        self.f2 = (w >> 16) & 0xffff
        self.c = (w >> 32) & 0xffff
        self.op = (w >> 48) & 15
        self.uses_p = self.op in (TERM, DIVACC, DIV, INITD)
        self.uses_s = (self.op == TERM and not (w & ONE) and not (w & INIT)) or self.op in (DIVACC, DIV, LOADS)


def live_ops(words, zero):
    """The micro-ops that do something: padding (a TERM on the constant +0.0 in a fast group, a NOP
    without a store in a slow one) is dropped -- the accumulator is dead wherever the compiler pads."""
    out = []
    for w in words:
        w = int(w)
        if (w & ~SLOW) == zero:
            continue
        o = Op(w)
        if o.op == NOP and not (w & (ST_SHARED | ST_HBM)):
            continue
        out.append(o)
    return out


def schedule_loads(ops, dist, chunk):
    """Where the HBM operand of each micro-op is requested: `dist` micro-ops ahead, but never before a
    store to the same element (read-after-write through memory) and never before its chunk starts."""
    last_store = {}
    for i, o in enumerate(ops):
        if o.uses_p:
            at = max(i - dist, last_store.get(o.f1, -1) + 1, (i // chunk) * chunk)
            o.load_at = min(at, i)
        else:
            o.load_at = -1
        if o.w & ST_HBM:
            last_store[o.c] = i


def emit_segment(name, words, zero, dist, chunk, out, div="div_inline", attr="__noinline__"):
    ops = live_ops(words, zero)
    schedule_loads(ops, dist, chunk)
    pending = {}  # position -> ops whose operand is requested there
    for i, o in enumerate(ops):
        o.name = f"p{i}"
        if o.uses_p:
            pending.setdefault(o.load_at, []).append(o)
    nchunks = (len(ops) + chunk - 1) // chunk
    for ck in range(nchunks):
        out.append(f"__device__ {attr} double2 {name}_{ck}(char *g, const unsigned sb, double t, double d)\n{{")
        for i in range(ck * chunk, min(len(ops), (ck + 1) * chunk)):
            for o in pending.get(i, ()):
                out.append(f'\tdouble {o.name}; asm volatile("ld.global.f64 %0, [%1+{o.f1 << 8}];" : "=d"({o.name}) : "l"(g) : "memory");')
            o = ops[i]
            w = o.w
            s = f"s{i}"
            if o.uses_s:
                out.append(f'\tdouble {s}; asm volatile("ld.shared.f64 %0, [%1+{o.f2 << 8}];" : "=d"({s}) : "r"(sb));')
            if o.op == TERM:
                if w & INIT:
                    out.append(f"\tt = {o.name};")
                else:
                    base = "0.0" if w & FIRSTZ else "t"
                    prod = o.name if w & ONE else f"{o.name} * {s}"
                    out.append(f"\tt = {base} - {prod};")
            elif o.op == DIVACC:
                out.append(f"\tt = {div}({s}, {o.name}); d = d - t * {s};")
            elif o.op == DIV:
                out.append(f"\tt = {div}({s}, {o.name});")
            elif o.op == INITD:
                out.append(f"\td = {o.name};")
            elif o.op == ENDD:
                out.append("\tt = regularise(d);")
            elif o.op == LOADS:
                out.append(f"\tt = {s};")
            if w & ST_HBM:
                out.append(f'\tasm volatile("st.global.f64 [%0+{o.c << 8}], %1;" :: "l"(g), "d"(t) : "memory");')
            if w & ST_SHARED:
                out.append(f'\tasm volatile("st.shared.f64 [%0+{o.c << 8}], %1;" :: "r"(sb), "d"(t) : "memory");')
        out.append("\treturn make_double2(t, d);\n}\n")
    # The warps of the CTA meet before every chunk, so that they fetch the same instructions at about the
    # same time (the code is far larger than the instruction caches); a warp with nothing to do skips the work.
    out.append(f"__device__ __forceinline__ void {name}(double *slice, const unsigned sb, const bool mine)\n{{")
    out.append("\tchar *g = reinterpret_cast<char *>(slice);\n\tdouble2 td = make_double2(0.0, 0.0);")
    for ck in range(nchunks):
        out.append(f"\t__syncthreads();\n\tif (mine) td = {name}_{ck}(g, sb, td.x, td.y);")
    out.append("}\n")
    return len(ops)


def generate(kind: str, dist: int = 32, chunk: int = 256, div: str = "div_inline", attr: str = "__noinline__") -> str:
    from . import dogbot, qpswift as Q
    batch = dogbot.generate(kind, 1)
    plan = Q.Plan(batch.pattern)
    info = plan.stream_info()
    assert info["ok"] == 1
    zero = info["layout"]["zero"]
    key = plan.stream_key()
    out = [f"// GENERATED by apf_quadruped_b200/codegen.py for the '{kind}' pattern (micro-program key {key:#018x}); do not edit.",
           f"// dist {dist} micro-ops of HBM operands in flight, chunks of {chunk} micro-ops.",
           '#include "stream.cuh"', '#include "stream_gen.cuh"', "", "namespace qpb {", f"namespace gen_{kind} {{", ""]
    counts = {}
    for seg in ("factor", "solve", "resid", "resz"):
        counts[seg] = emit_segment(seg, plan.stream_words(seg), zero, dist, chunk, out, div, attr)
    out += ["struct Prog", "{", "\tstatic constexpr bool cta_sync = true;",
            "\tstatic __device__ __forceinline__ void factor(const StreamView &, double *slice, const unsigned sb, const bool mine) { gen_%s::factor(slice, sb, mine); }" % kind,
            "\tstatic __device__ __forceinline__ void solve(const StreamView &, double *slice, const unsigned sb, const bool mine) { gen_%s::solve(slice, sb, mine); }" % kind,
            "\tstatic __device__ __forceinline__ void resid(const StreamView &, double *slice, const unsigned sb, const bool mine) { gen_%s::resid(slice, sb, mine); }" % kind,
            "\tstatic __device__ __forceinline__ void resz(const StreamView &, double *slice, const unsigned sb, const bool mine) { gen_%s::resz(slice, sb, mine); }" % kind,
            "};", "", f"}} // namespace gen_{kind}", "",
            f"QPB_REGISTER_GENERATED({kind}, {key:#018x}ull, gen_{kind}::Prog)", "", "} // namespace qpb", ""]
    plan.close()
    return "\n".join(out), counts


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("kinds", nargs="*", default=["trot", "stance", "crawl"])
    ap.add_argument("--out", default="build/gen")
    ap.add_argument("--dist", type=int, default=32)
    ap.add_argument("--chunk", type=int, default=256)
    ap.add_argument("--div", default="div_inline", help="div_rn: out-of-line division; div_inline: inlined")
    ap.add_argument("--attr", default="__noinline__")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for kind in a.kinds:
        src, counts = generate(kind, a.dist, a.chunk, a.div, a.attr)
        path = os.path.join(a.out, f"stream_gen_{kind}.cu")
        old = open(path).read() if os.path.exists(path) else None
        if old != src:  # keep the time stamp when nothing changed (make)
            with open(path, "w") as f:
                f.write(src)
        print(f"{path}: {counts}")


if __name__ == "__main__":
    main()
