import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from oracle import so3 as oso3
from gpr_calculator_b200.SO3 import SO3
from gpr_calculator_b200.utilities import SimpleAtoms

def cu_fcc(nrep, a=3.61, seed=0, noise=0.05):
    base = np.array([[0,0,0],[0.5,0.5,0],[0.5,0,0.5],[0,0.5,0.5]])*a
    pos = np.concatenate([base + np.array([i,j,k])*a for i in range(nrep) for j in range(nrep) for k in range(nrep)])
    rng = np.random.default_rng(seed)
    return SimpleAtoms([29]*len(pos), pos + rng.normal(scale=noise, size=pos.shape), np.eye(3)*a*nrep)

for at, prm in ((cu_fcc(2, seed=2000), (3,4,5.0,2.0)), (cu_fcc(3, seed=1000), (3,4,5.0,2.0))):
    des = SO3(nmax=prm[0], lmax=prm[1], rcut=prm[2], alpha=prm[3])
    t = time.time(); r = des.calculate(at); torch.cuda.synchronize(); t1 = time.time()-t
    t = time.time(); r = des.calculate(at); torch.cuda.synchronize(); t1 = time.time()-t
    t = time.time(); x, dxdr, seq = oso3.so3_calculate(at.positions, at.cell, at.pbc, at.numbers, *prm); t2 = time.time()-t
    print(len(at), 'gpu s', t1, 'oracle s', t2, 'seq equal', np.array_equal(r['seq'], seq), r['seq'].shape,
          'x', np.abs(r['x']-x).max()/np.abs(x).max(), 'dxdr', np.abs(r['dxdr']-dxdr).max()/np.abs(dxdr).max())
rng = np.random.default_rng(5)
pos = rng.uniform(0, 6, size=(13, 3))
at2 = SimpleAtoms([13]*12+[79], pos, np.diag([5.73, 5.73, 13.75]), pbc=(True, True, False))
for prm in ((3,4,5.0,2.0), (4,3,4.0,1.5), (2,6,3.5,2.0), (1,0,3.0,1.0)):
    des = SO3(nmax=prm[0], lmax=prm[1], rcut=prm[2], alpha=prm[3])
    r = des.calculate(at2)
    x, dxdr, seq = oso3.so3_calculate(at2.positions, at2.cell, at2.pbc, at2.numbers, *prm)
    print(prm, 'seq equal', np.array_equal(r['seq'], seq), 'x', np.abs(r['x']-x).max()/np.abs(x).max(), 'dxdr', np.abs(r['dxdr']-dxdr).max()/max(np.abs(dxdr).max(),1e-300))
# isolated atom / molecule in big box
at3 = SimpleAtoms([1, 1, 8], [[0,0,0],[0,0,0.96],[8,8,8]], np.eye(3)*20, pbc=(False,False,False))
des = SO3(nmax=3, lmax=4, rcut=5.0)
r = des.calculate(at3); x, dxdr, seq = oso3.so3_calculate(at3.positions, at3.cell, at3.pbc, at3.numbers, 3,4,5.0,2.0)
print('mol seq', r['seq'].tolist(), seq.tolist(), np.abs(r['x']-x).max(), np.abs(r['dxdr']-dxdr).max())
# batch
ats = [cu_fcc(2, seed=3000+k) for k in range(64)]
t=time.time(); out = des.calculate_batch(ats, to_host=False); torch.cuda.synchronize(); print('batch64 s', time.time()-t, out['x'].shape, out['dxdr'].shape)
t=time.time(); out = des.calculate_batch(ats, to_host=False); torch.cuda.synchronize(); print('batch64 s', time.time()-t)
