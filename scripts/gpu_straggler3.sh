for envs in "X=1" "QPB_STREAM_SERVER=0"; do
echo "=== $envs"
env $envs timeout 300 python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from apf_quadruped_b200 import dogbot, qpswift as Q
for cnt in (50000, 37000):
    b = dogbot.generate('trot', cnt, start=4 * 65536)
    s = Q.BatchSolver(b.pattern, b.count)
    best = 1e9
    for rep in range(3):
        out = s.solve(b); best = min(best, s.last_kernel_ms)
    bad = np.nonzero(out['flag'])[0]
    print('RESULT count %d: %.2f ms, stragglers at %s' % (cnt, best, list(bad)), flush=True)
    s.close()
PY
done
