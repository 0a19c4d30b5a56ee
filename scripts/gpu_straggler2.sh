for envs in "QPB_STREAM_SPILL=22" "QPB_STREAM_SPILL=24" "QPB_STREAM_SPILL=28" "QPB_STREAM_SPILL=22 QPB_STREAM_SERVER=0"; do
echo "=== $envs"
env $envs timeout 300 python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from apf_quadruped_b200 import dogbot, qpswift as Q
b = dogbot.generate('trot', 65536, start=4 * 65536)
s = Q.BatchSolver(b.pattern, b.count)
best = 1e9
for rep in range(3):
    out = s.solve(b); best = min(best, s.last_kernel_ms)
bad = np.nonzero(out['flag'])[0]
print('RESULT %.2f ms, stragglers at %s' % (best, list(bad)), flush=True)
s.close()
PY
done
