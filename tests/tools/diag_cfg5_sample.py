import sys, math, numpy as np
import pathlib; R=pathlib.Path(__file__).resolve().parents[2]; sys.path.insert(0,str(R)); sys.path.insert(0,str(R/'tests'))
import os_stab_b200 as osb, oracle_binding as ob
import bench
o = ob.load()
eng = osb.Engine(0)
prof = eng.solve_bl()
oprof,_,_ = o.solve_bl()
opts = osb.shoot_defaults(osb.FLOW_CONTEBL)
idx, nre = bench.workload_indices(1, 0)
rng = np.random.default_rng(1)
sub = np.sort(rng.choice(idx.size, 512, replace=False))
al, Re, c0 = bench.build_inputs(idx[sub], nre)
r = eng.shoot_temporal(opts, prof, al, Re, c0, want_resid=True, hist_cap=20)
ref = o.shoot_temporal_batch(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, oprof, al, Re, c0)
print('gpu status hist', np.bincount(r['status'], minlength=3), 'ref', np.bincount(ref['status'], minlength=3))
print('gpu passes hist', np.bincount(r['passes'], minlength=21))
print('ref passes hist', np.bincount(ref['passes'], minlength=21))
ok = (ref['status']==0)&(r['status']==0)
rel = np.abs(r['c']-ref['c'])/np.abs(ref['c'])
print('max rel (both conv)', rel[ok].max(), 'all', rel.max())
bad = np.where(r['status']!=0)[0][:5]
for b in bad:
    print('point', b, 'alpha', al[b].real/math.sqrt(2), 'Re', Re[b]/math.sqrt(2), 'seed', c0[b])
    rr = o.shoot_temporal(ob.FLOW_CONTEBL, ob.INT_CK45, 1000, 10.0, 0.0, 17.0, oprof, al[b], Re[b], c0[b], hist_cap=20)
    print(' ref passes', rr['passes'], 'c', rr['c'])
    for k in range(min(20, r['passes'][b])):
        h = r['history'][b][k]
        hr = rr['history'][k] if k < rr['passes'] else [0,0,0,0]
        print('  %2d gpu c=%.15f%+.15fi err=%.3e%+.3ei | ref c=%.15f%+.15fi err=%.3e%+.3ei' % (k+1, h[0],h[1],h[2],h[3], hr[0],hr[1],hr[2],hr[3]))
