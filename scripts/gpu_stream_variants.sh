# timing of the throughput kernel under env-var variants: bash scripts/gpu_stream_variants.sh "A=1 B=2" "C=3" ...
KINDS=${KINDS:-trot}
for envs in "$@"; do
  echo "=== $envs"
  env $envs timeout 300 python - <<'PY'
import time, numpy as np, sys, os
sys.path.insert(0,'.')
from apf_quadruped_b200 import dogbot, qpswift as Q
for kind in os.environ.get('KINDS','trot').split(','):
    cnt = int(os.environ.get('COUNT', 65536 if kind != 'mpc' else 16384))
    b = dogbot.generate(kind, cnt)
    s = Q.BatchSolver(b.pattern, cnt)
    best = 1e9
    for rep in range(3):
        out = s.solve(b); best = min(best, s.last_kernel_ms)
    print('RESULT', kind, cnt, s.last_kernel, 'generated' if s.last_kernel_generated else 'interpreted', '%.0f solves/s' % (cnt / best * 1e3), 'iters', out['iters'].sum(), 'flags', np.bincount(out['flag']), flush=True)
    s.close()
PY
done
