"""Golden vectors of the descriptor options outside the default path (SO3(weight_on=True), calculate(atom_ids=...),
derivative=False), produced by the UNMODIFIED reference under the stubs of oracle/ref_harness.py (build container
only):  python tests/golden/gen_golden_so3_options.py   ->  so3_options.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from oracle import ref_harness as rh   # noqa: E402
from gen_golden import slab            # noqa: E402


def main():
    os.chdir("/tmp")
    m = rh.modules()
    out = {}
    strucs = [slab(11, n_fixed=0),
              rh.Atoms([1, 1, 16, 46, 46], [[0, 0, 0.3], [0.2, 1.5, 0.1], [0.9, 0.7, 0.9], [3.0, 3.0, 3.0], [5.5, 3.1, 2.7]],
                       [[8.0, 0.5, 0], [0, 8.0, 0], [0, 0, 9.0]], pbc=(True, True, True))]
    prms = [(3, 4, 5.0, 2.0), (2, 3, 4.0, 1.5)]
    ids = [[12, 3, 7], [4, 0]]
    for k, (at, prm, sub) in enumerate(zip(strucs, prms, ids)):
        out["s%d_numbers" % k], out["s%d_pos" % k], out["s%d_cell" % k], out["s%d_pbc" % k] = at.numbers, at.positions, np.asarray(at.cell), at.pbc
        out["s%d_prm" % k], out["s%d_ids" % k] = np.array(prm), np.array(sub)
        kw = dict(nmax=prm[0], lmax=prm[1], rcut=prm[2], alpha=prm[3])
        r = m.SO3(weight_on=True, **kw).calculate(at)
        out["s%d_w_x" % k], out["s%d_w_dxdr" % k], out["s%d_w_seq" % k] = r["x"], r["dxdr"], r["seq"]
        r = m.SO3(**kw).calculate(at, atom_ids=sub)
        out["s%d_sub_x" % k], out["s%d_sub_dxdr" % k], out["s%d_sub_seq" % k] = r["x"], r["dxdr"], r["seq"]
        r = m.SO3(derivative=False, weight_on=True, **kw).calculate(at)
        out["s%d_nod_x" % k] = r["x"]
        assert r["dxdr"] is None
    np.savez_compressed(os.path.join(HERE, "so3_options.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
