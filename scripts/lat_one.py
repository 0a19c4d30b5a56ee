"""The throughput kernel's straggler through the latency kernel: trot instance 264420 (maxit = 100), batch of one."""
import os, sys
sys.path.insert(0, '.')
os.environ['QPB_STREAM_MIN'] = '100000000'
from apf_quadruped_b200 import dogbot, qpswift as Q
b = dogbot.generate('trot', 1, start=264420)
s = Q.BatchSolver(b.pattern, 1)
for rep in range(3):
    out = s.solve(b)
print('kernel %.3f ms iterations %d' % (s.last_kernel_ms, out['iters'][0]))
