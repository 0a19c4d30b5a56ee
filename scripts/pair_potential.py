"""How many HBM requests of the generated programs a pair-interleaved layout would save (round-2 lever):
elements 2k and 2k+1 of a lane adjacent, so that one 16-byte request serves both when both are loaded within
`window` micro-ops of each other (a 16-byte request costs a warp ~15 % more than an 8-byte one,
scripts/micro/randline3.cu).  Counts loads only (stores pair the same way); current element numbering."""
import sys
sys.path.insert(0, '.')
from apf_quadruped_b200 import codegen, dogbot, qpswift as Q

for kind in ("trot", "stance", "crawl"):
    plan = Q.Plan(dogbot.generate(kind, 1).pattern)
    segs, _ = codegen.build_all(plan)
    print(kind)
    for window in (8, 32):
        total = paired = 0
        line = []
        for seg, B in segs.items():
            loads = saved = 0
            for fn in B.functions:
                recs = sorted(fn.loads, key=lambda r: r["use"])
                taken = set()
                for i, r in enumerate(recs):
                    loads += 1
                    if i in taken:
                        continue
                    for j in range(i + 1, len(recs)):
                        if recs[j]["use"] - r["use"] > window:
                            break
                        if j not in taken and recs[j]["e"] == (r["e"] ^ 1):
                            taken.add(j)
                            saved += 1
                            break
            line.append(f"{seg} {loads}->{loads - saved}")
            total += loads
            paired += saved
        print(f"  window {window:2d}: " + ", ".join(line) + f"; all {total}->{total - paired} (-{100 * paired / total:.0f} %)")
    plan.close()
