# parity of the throughput kernel + timing of both kernels on the bench workloads
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_stream_gpu.py -x -q 2>&1 | tail -15
timeout 900 python - <<'PY'
import time, numpy as np, sys, os
sys.path.insert(0,'.')
from apf_quadruped_b200 import dogbot, qpswift as Q
for kind, cnt in [('trot', 65536), ('stance', 65536), ('crawl', 32768), ('mpc', 16384)]:
    b = dogbot.generate(kind, cnt)
    s = Q.BatchSolver(b.pattern, cnt)
    for rep in range(3):
        t = time.time(); out = s.solve(b); dt = time.time() - t
        print(kind, cnt, s.last_kernel, 'e2e %.1f ms, kernel %.2f ms -> %.0f solves/s kernel; iters mean %.2f max %d; flags %s' % (dt*1e3, s.last_kernel_ms, cnt / s.last_kernel_ms * 1e3, out['iters'].mean(), out['iters'].max(), np.bincount(out['flag'])), flush=True)
    s.close()
PY
