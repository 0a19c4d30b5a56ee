# the shard that owns the maxit straggler (trot instance 264420 = QP 2276 of shard 4) against a shard without one
timeout 600 python -m pytest tests/test_stream_gpu.py -x -q 2>&1 | tail -3
for envs in "X=1" "QPB_STREAM_SERVER=0" "QPB_STREAM_SERVER=0 QPB_STREAM_SPILL=32"; do
echo "=== $envs"
env $envs timeout 300 python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from apf_quadruped_b200 import dogbot, qpswift as Q
for start in (0, 4 * 65536):
    b = dogbot.generate('trot', 65536, start=start)
    s = Q.BatchSolver(b.pattern, b.count)
    best = 1e9
    for rep in range(3):
        out = s.solve(b); best = min(best, s.last_kernel_ms)
    print('RESULT shard start %d: %.2f ms (%d launches), iterations max %d, flags %s' % (start, best, s.last_launches, out['iters'].max(), np.bincount(out['flag'])), flush=True)
    s.close()
PY
done
