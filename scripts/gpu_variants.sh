# time trot (and optionally other kinds) for several env-var / library variants
set -x
KINDS=${KINDS:-trot}
for lib in apf_quadruped_b200/lib/variants/*.so; do
  for envs in "$@"; do
    echo "=== $(basename $lib) $envs"
    env $envs QPB_LIB=$lib timeout 300 python - <<PY
import time, numpy as np, sys, os
sys.path.insert(0,'.')
from apf_quadruped_b200 import dogbot, qpswift as Q
for kind in os.environ.get('KINDS','trot').split(','):
    cnt = 32768 if kind != 'mpc' else 8192
    b = dogbot.generate(kind, cnt)
    s = Q.BatchSolver(b.pattern, cnt)
    best = 1e9
    for rep in range(3):
        out = s.solve(b); best = min(best, s.last_kernel_ms)
    print('RESULT', os.path.basename(os.environ['QPB_LIB']), os.environ.get('TAG',''), kind, '%.0f solves/s' % (cnt / best * 1e3), 'iters', out['iters'].sum(), flush=True)
    s.close()
PY
  done
done
