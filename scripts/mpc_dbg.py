import sys, numpy as np
sys.path.insert(0,'.')
from apf_quadruped_b200 import dogbot, qpswift as Q
b = dogbot.generate('mpc', 300)
s = Q.BatchSolver(b.pattern, 300)
out = s.solve(b)
print('ok', out['iters'].mean(), np.bincount(out['flag']))
