import sys, os, numpy as np
sys.path.insert(0,'.')
from apf_quadruped_b200 import dogbot, qpswift as Q
names=['resid+setup','cta wait','factor','solve1','mid','solve2','update','loop']
for kind in os.environ.get('KINDS','trot').split(','):
    cnt = 32768 if kind != 'mpc' else 8192
    b = dogbot.generate(kind, cnt)
    s = Q.BatchSolver(b.pattern, cnt)
    out = s.solve(b); out = s.solve(b)
    t = s.read_timing(); kt = t[8:]; t = t[:8]; tot = sum(t)
    its = out['iters'].sum() + cnt
    print(kind, '%.0f solves/s'%(cnt/s.last_kernel_ms*1e3), 'cycles per QP-iteration:', ' | '.join('%s %.0f (%.0f%%)'%(n, v/its, 100*v/tot) for n,v in zip(names,t)), flush=True)
    print('    per super-wave:', ' | '.join('%s %.0f cycles x %.0f/iter' % (n, kt[i] / max(1, kt[4 + i]), kt[4 + i] / its) for i, n in enumerate(['mul', 'div', 'sum', 'prod'])), flush=True)
    s.close()
