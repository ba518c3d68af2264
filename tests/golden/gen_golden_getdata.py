"""Golden vectors of the database -> training-rows helpers (utilities.get_data / convert_struc / get_strucs), produced by
the UNMODIFIED reference under the stubs of oracle/ref_harness.py (build container only):

    python tests/golden/gen_golden_getdata.py   ->  getdata.db (a 3-row ASE sqlite file) + getdata.npz

The reference opens the file with ase.db.connect (absent here); a stand-in built on the product's sqlite reader serves the
rows, everything after that (descriptor, row selection, labels) is the reference's own code.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from oracle import ref_harness as rh          # noqa: E402
from gen_golden import slab, toy_labels       # noqa: E402
from gpr_calculator_b200 import asedb         # noqa: E402


class _Conn:
    def __init__(self, filename):
        self.rows = list(asedb.read_rows(filename))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def select(self):
        for r in self.rows:
            yield types.SimpleNamespace(id=r.id, data=types.SimpleNamespace(keys=lambda d=r.data: d.keys(), **r.data))

    def get_atoms(self, id):
        r = [r for r in self.rows if r.id == id][0]
        return rh.Atoms(r.numbers, r.positions, r.cell, r.pbc)


def main():
    os.chdir("/tmp")
    m = rh.modules()
    db = os.path.join(HERE, "getdata.db")
    strucs = [slab(21 + k, n_fixed=0) for k in range(3)]
    labels = [toy_labels(s, 40 + k) for k, s in enumerate(strucs)]
    asedb.write_rows(db, [(s, {"tag": k}, {"energy": E, "force": F}) for k, (s, (E, F)) in enumerate(zip(strucs, labels))])
    m.utilities.connect = lambda filename, serial=True: _Conn(filename)
    des = m.SO3(nmax=3, lmax=4, rcut=5.0)
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        cases = {"all": m.utilities.get_data(db, des), "cap": m.utilities.get_data(db, des, N_force=17),
                 "sel": m.utilities.get_data(db, des, lists=[0, 2], select=True), "noe": m.utilities.get_data(db, des, N_force=5, no_energy=True)}
    for name, data in cases.items():
        out[name + "_nE"], out[name + "_nF"] = len(data["energy"]), len(data["force"])
        if data["energy"]:
            out[name + "_E_x"] = np.concatenate([x for x, _, _ in data["energy"]])
            out[name + "_E_y"] = np.array([y for _, y, _ in data["energy"]])
            out[name + "_E_ele"] = np.concatenate([e for _, _, e in data["energy"]])
        out[name + "_F_rows"] = np.array([len(x) for x, _, _, _ in data["force"]])
        if name == "all":        # the full set: shapes, labels and checksums only (keeps the fixture small)
            out["all_F_x_abs_sum"] = np.abs(np.concatenate([x for x, _, _, _ in data["force"]])).sum()
            out["all_F_dxdr_abs_sum"] = np.abs(np.concatenate([d for _, d, _, _ in data["force"]])).sum()
            out["all_F_y"] = np.array([y for _, _, y, _ in data["force"]])
            out["all_db_fids"] = np.array([len(f) for _, _, _, _, f in data["db"]])
            continue
        out[name + "_F_x"] = np.concatenate([x for x, _, _, _ in data["force"]])
        out[name + "_F_dxdr"] = np.concatenate([d for _, d, _, _ in data["force"]])
        out[name + "_F_y"] = np.array([y for _, _, y, _ in data["force"]])
        out[name + "_F_ele"] = np.concatenate([e for _, _, _, e in data["force"]])
        out[name + "_db_fids"] = np.array([len(f) for _, _, _, _, f in data["db"]])
        out[name + "_db_E"] = np.array([E for _, E, _, _, _ in data["db"]])
    S, V = m.utilities.get_strucs(db, N_max=2)
    out["strucs_n"] = len(S)
    out["strucs_E"] = np.array([v[0] for v in V])
    np.savez_compressed(os.path.join(HERE, "getdata.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items() if k.startswith(("cap", "strucs"))})


if __name__ == "__main__":
    main()
