"""Small run of every kernel for compute-sanitizer (memcheck): latency kernel, generated and interpreted throughput
kernels incl. stragglers' server and hand-over, dense front-end."""
import os, sys
sys.path.insert(0, '.')
os.environ['QPB_STREAM_MIN'] = '1'
import numpy as np
from apf_quadruped_b200 import dogbot, qpswift as Q
b = dogbot.generate('trot', 300, start=264320)        # contains the maxit instance
s = Q.BatchSolver(b.pattern, b.count); o = s.solve(b); print('generated', s.last_kernel_generated, o['iters'].max()); s.close()
os.environ['QPB_STREAM_INTERPRET'] = '1'
s = Q.BatchSolver(b.pattern, b.count); o2 = s.solve(b); print('interpreted', not s.last_kernel_generated, all(np.array_equal(o[k], o2[k]) for k in ('x', 'z', 'flag', 'iters'))); s.close()
del os.environ['QPB_STREAM_INTERPRET']
c = dogbot.generate('crawl', 97)
s = Q.BatchSolver(c.pattern, c.count); P, A, G = c.dense_batch(); o3 = s.solve_dense(P, A, G, c.c, c.h, c.b); print('dense', o3['violation'].any(), o3['iters'].max()); s.close()
os.environ['QPB_STREAM_MIN'] = '1000000'
d = dogbot.generate('stance', 40)
s = Q.BatchSolver(d.pattern, d.count); o4 = s.solve(d); print('latency', s.last_kernel, o4['iters'].max()); s.close()
