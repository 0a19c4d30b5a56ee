# latency kernel: the single maxit trot instance (the throughput kernel's straggler), a small batch, and a full batch
python - <<'PY'
import sys, time, os, numpy as np
sys.path.insert(0, '.')
from apf_quadruped_b200 import dogbot, qpswift as Q
os.environ['QPB_STREAM_MIN'] = '100000000'
for start, cnt in ((264420, 1), (264320, 256), (0, 1036), (0, 32768)):
    b = dogbot.generate('trot', cnt, start=start)
    s = Q.BatchSolver(b.pattern, cnt)
    best = 1e9
    for rep in range(3):
        out = s.solve(b); best = min(best, s.last_kernel_ms)
    print('LAT trot start %d count %d: %s kernel %.3f ms, %.0f solves/s, iterations total %d max %d' % (start, cnt, s.last_kernel, best, cnt / best * 1e3, out['iters'].sum(), out['iters'].max()), flush=True)
    s.close()
PY
