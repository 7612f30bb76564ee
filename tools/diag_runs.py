import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/model-predictive-control-tuning_b200')
import mpcgpu
import os
p = mpcgpu.shell7x5() if os.environ.get('DIAG_CASE') == 'shell7x5' else mpcgpu.shell3x3(2)
NIT = p.nit
ev = mpcgpu.Evaluator(p, device=0)
NPOP = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N, Nu, dl, lm = mpcgpu.synthetic_population(p, NPOP, seed=0)
ev.lib.mpcgpu_debug_enable_diag.argtypes=[C.c_void_p, C.c_int]; ev.lib.mpcgpu_debug_get_diag.argtypes=[C.c_void_p, C.c_void_p]
ev.lib.mpcgpu_debug_enable_diag(ev.h, 1)
for _ in range(2): out = ev.eval_batch(N, Nu, dl, lm, mode='gam')
d = np.zeros((NPOP,4), dtype=np.uint64)
rc = ev.lib.mpcgpu_debug_get_diag(ev.h, d.ctypes.data_as(C.c_void_p)); print('rc', rc)
cyc = d[:,3].astype(float); its = d[:,1].astype(float); con = d[:,0].astype(float); qm = (d[:,2] & np.uint64(0xffffffff)).astype(float); xc = (d[:,2] >> np.uint64(32)).astype(float); print('re-entries: total %d, runs with any %d' % (xc.sum(), (xc > 0).sum()))
print('cycles: mean %.3g median %.3g p99 %.3g max %.3g (ms at 1.965GHz: max %.2f)'%(cyc.mean(), np.median(cyc), np.percentile(cyc,99), cyc.max(), cyc.max()/1.965e6))
print('counters', ev.counters()); print('sum of run cycles %.3g -> /(148 SMs x 3 CTAs) = %.1f ms; heaviest run %.1f ms' % (cyc.sum(), cyc.sum()/444/1.965e6, cyc.max()/1.965e6))
top = np.argsort(-cyc)[:12]
for c in top: print(c, 'N',N[c],'Nu',Nu[c],'cyc %.3g its %d con %d qmax %d'%(cyc[c], its[c], con[c], qm[c]), 'cyc/step %.0f'%(cyc[c]/NIT), 'lam', lm[c])
# fast candidates
fast = np.where(con==0)[0]
for P,(lo,hi) in {4:(1,4),8:(5,8),16:(9,15)}.items():
    sel = fast[(Nu[fast]>=lo)&(Nu[fast]<=hi)]
    if len(sel): print('P',P,'fast-only candidates', len(sel), 'cycles/step median %.0f'%(np.median(cyc[sel])/NIT))
# regression cycles vs its, con
A = np.stack([np.ones(NPOP), con, its, its*qm],1); coef,*_ = np.linalg.lstsq(A, cyc, rcond=None); print('fit cyc ~ %.3g + %.3g*con + %.3g*its + %.3g*its*qmax'%tuple(coef))
np.save('gpurun_out/diag_%s.npy' % (sys.argv[2] if len(sys.argv) > 2 else 'x'), d)
