import sys, time, json, numpy as np
sys.path.insert(0, '/root/repo')
from tests import common
from oracle import pvoracle
from planeverb_b200 import pvcuda
scenes = common.load_scenes()
print('devices', pvcuda.device_count())
def check(scene, n=None, res=275, T=0, sk=0, variant=0, nl=1):
    if n is None: size, scale = 25.0, 1.0
    else: size, scale = common.scaled_config(n, res)
    O = pvoracle.OracleSim(size, size, res, T=T)
    G = pvcuda.Scene(size, size, res, T=T, max_sources=nl, step_kernel=sk, variant=variant)
    print(scene, 'n', G.gx, 'T', G.T, 'sk', sk, 'var', variant, 'efree', O.efree, G.efree, O.efree == G.efree)
    for b in common.boxes_of(scenes, scene, scale):
        O.add_aabb(*b); G.add_aabb(*b)
    ob, oR = O.coef(); gb, gy = G.coef()
    assert np.array_equal(ob, gb)
    Ls = common.listeners_for(nl, scale)
    res_g, dly_g = G.solve(Ls)
    print('   timing', G.timing())
    for i, L in enumerate(Ls):
        O.generate(L, keep_velocity=True); O.analyze(L)
        for t in (0, 1, 3, 4, 7, 50, G.T-1):
            gp = G.pressure(t, i); op = O.hist[t].reshape(gp.shape)
            if not common.bit_equal(gp, op).all():
                bad = ~common.bit_equal(gp, op)
                print('   PRESSURE MISMATCH t', t, 'count', bad.sum(), 'first', np.argwhere(bad)[:5], np.abs(gp-op).max()); break
        else: print('   pressure planes bit-equal')
        ir = G.ir(O.gx//2, O.gy//3, i); idx = (O.gx//2)*(O.gy+1) + O.gy//3
        print('   ir p/vx/vy equal', common.bit_equal(ir[:,0], O.hist[:, idx]).all(), common.bit_equal(ir[:,1], O.hvx[:, idx]).all(), common.bit_equal(ir[:,2], O.hvy[:, idx]).all())
        print('   delay equal', np.array_equal(dly_g[i], O.delay))
        valid = (O.delay < 3e38) & (O.clamped == 0)
        for k, name in enumerate(common.FIELDS):
            m = valid if k not in (4,5) else np.ones_like(valid)
            be = common.bit_equal(res_g[i][m, k], O.results[m, k])
            print('    ', name, 'bit-equal', be.all(), 'nmis', (~be).sum(), 'maxrel', common.rel_err(res_g[i][m,k], O.results[m,k]).max() if m.any() else 0)
    G.close()
for sk in (1, 0):
    check('SmallRoom', sk=sk)
    check('FloorPlanScene', sk=sk)
    check('Shoebox', n=128, T=500, sk=sk, nl=2)
for v in (1,7,8,9):
    check('BigRoom', n=200, T=300, variant=v)
# timing at 1024
size, scale = common.scaled_config(1024)
for sk, var in ((0,0),(0,7),(0,8),(0,9),(0,5),(0,6)):
    G = pvcuda.Scene(size, size, 275, T=1000, max_sources=4, step_kernel=sk, variant=var, efree=0.0447895788)
    for b in common.boxes_of(scenes, 'BigRoom', scale): G.add_aabb(*b)
    Ls = common.listeners_for(4, scale)
    G.solve(Ls, fetch=False); G.solve(Ls, fetch=False)
    st, an, tot, nl = G.timing()
    cu = 1024*1024*1000*4
    print(f'1024^2 x4 T=1000 sk={sk} var={var}: steps {st:.1f} ms ({cu/st/1e6:.1f} Gcell/s) analyzer {an:.1f} ms total {tot:.1f} ms ({cu/tot/1e6:.1f} Gcell/s) launches {nl}')
    G.close()
