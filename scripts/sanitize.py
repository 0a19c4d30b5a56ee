"""Small run of every kernel for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python scripts/sanitize.py
generated warp-per-QP kernels (value row in the global scratch row) for the three DogBot patterns incl. a maxit instance,
the single-QP path, the interpreted warp-per-QP kernel (MPC), generated and interpreted thread-per-QP kernels incl. the
stragglers' server and hand-over, the dense front-end, the WBC formulation / torque kernels, the indexed solve and a few
rollout ticks (eager and from a CUDA graph)."""
import os
import sys

sys.path.insert(0, '.')
import numpy as np
import torch

from apf_quadruped_b200 import dogbot, qpswift as Q, rollout, wbc

for kind, cnt, start in (('trot', 300, 264320), ('stance', 70, 0), ('crawl', 97, 0)):  # (trot: contains the maxit instance)
    b = dogbot.generate(kind, cnt, start=start)
    s = Q.BatchSolver(b.pattern, b.count)
    o = s.solve(b)
    print('warp', kind, s.last_kernel, s.last_kernel_generated, o['iters'].max())
    s.close()
one = dogbot.generate('stance', 1)
print('single', Q.solve_single(one.pattern, one.Ppr[0], one.Apr[0], one.Gpr[0], one.c[0], one.h[0], one.b[0])['iters'])
m = dogbot.generate('mpc', 12)
s = Q.BatchSolver(m.pattern, m.count); o = s.solve(m); print('mpc', s.last_kernel, s.last_kernel_generated, o['iters'].max()); s.close()
os.environ['QPB_STREAM_MIN'] = '1'
b = dogbot.generate('trot', 300, start=264320)
s = Q.BatchSolver(b.pattern, b.count); o = s.solve(b); print('stream generated', s.last_kernel, s.last_kernel_generated, o['iters'].max()); s.close()
os.environ['QPB_STREAM_INTERPRET'] = '1'
s = Q.BatchSolver(b.pattern, b.count); o2 = s.solve(b); print('stream interpreted', not s.last_kernel_generated, all(np.array_equal(o[k], o2[k]) for k in ('x', 'z', 'flag', 'iters'))); s.close()
del os.environ['QPB_STREAM_INTERPRET'], os.environ['QPB_STREAM_MIN']
c = dogbot.generate('crawl', 97)
s = Q.BatchSolver(c.pattern, c.count); P, A, G = c.dense_batch(); o3 = s.solve_dense(P, A, G, c.c, c.h, c.b); print('dense', o3['violation'].any(), o3['iters'].max()); s.close()
# WBC formulation + torques, plain and through an index list
dev = torch.device('cuda', 0)
S = wbc.random_states(50)
pat = dogbot.pattern('crawl')
s = Q.BatchSolver(pat, 50)
f = wbc.Formulator(s, 'crawl')
dS = torch.from_numpy(S).to(dev)
rows = [torch.zeros((50, w), dtype=torch.float64, device=dev) for w in (pat.nnzP, pat.nnzA, pat.nnzG, pat.n, pat.m, pat.p)]
f.formulate(50, dS.data_ptr(), *[r.data_ptr() for r in rows])
idx = torch.tensor([3, 49, 0, 17], dtype=torch.int32, device=dev)
cnt = torch.tensor([4], dtype=torch.int32, device=dev)
out = [torch.zeros((50, w), dtype=torch.float64, device=dev) for w in (pat.n, pat.p, pat.m, pat.m)]
fl, it = torch.zeros(50, dtype=torch.int32, device=dev), torch.zeros(50, dtype=torch.int32, device=dev)
s.solve_device_indexed(50, Q.default_settings(), idx.data_ptr(), cnt.data_ptr(), *[r.data_ptr() for r in rows], *[o.data_ptr() for o in out], fl.data_ptr(), it.data_ptr(), None)
tau = torch.zeros((50, 12), dtype=torch.float64, device=dev)
f.torques(50, dS.data_ptr(), out[0].data_ptr(), tau.data_ptr(), dindex=idx.data_ptr(), dcount=cnt.data_ptr())
torch.cuda.synchronize()
print('wbc', it.cpu().numpy()[[3, 49, 0, 17]], float(tau.abs().max()))
s.close()
for graph in (False, True):
    ro = rollout.Rollout(torch, 40, device=0, graph=graph)
    ro.step(6)
    snap = ro.snapshot()
    print('rollout graph' if graph else 'rollout eager', snap['tick'], snap['iters'].max(), int((snap['flag'] != 0).sum()))
    ro.close()
