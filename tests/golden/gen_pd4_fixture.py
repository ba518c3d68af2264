"""Convert the reference's only real-data fixture, examples/database/pd4-RBF.db (206 labelled Pd4/MgO structures
written by GP.export_ase_db), into a compact npz for the tests (the GPU box has no /root/reference):

    python tests/golden/gen_pd4_fixture.py      # build container only

pd4.npz: numbers [220], cell [3,3], pbc [3], positions [206,220,3], energy [206], force [206,220,3],
energy_in [206], force_in (flat) + force_in_ptr [207], dft_energy [206], plus the model file pd4-RBF.json
(hyper-parameters, noise, descriptor settings and the stored error block) as a JSON string.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from gpr_calculator_b200 import asedb   # noqa: E402

REF = os.environ.get("GPR_REFERENCE_ROOT", "/root/reference")


def main():
    rows = list(asedb.read_rows(os.path.join(REF, "examples", "database", "pd4-RBF.db")))
    assert all(np.array_equal(r.numbers, rows[0].numbers) for r in rows)
    fin = [np.asarray(r.data["force_in"], dtype=np.int64) for r in rows]
    out = {
        "numbers": rows[0].numbers, "cell": rows[0].cell, "pbc": rows[0].pbc,
        "positions": np.stack([r.positions for r in rows]),
        "energy": np.array([r.data["energy"] for r in rows]),
        "force": np.stack([np.asarray(r.data["force"]) for r in rows]),
        "energy_in": np.array([bool(r.data["energy_in"]) for r in rows]),
        "force_in": np.concatenate(fin), "force_in_ptr": np.concatenate(([0], np.cumsum([len(f) for f in fin]))),
        "dft_energy": np.array([r.key_value_pairs["dft_energy"] for r in rows]),
        "model_json": np.array(open(os.path.join(REF, "examples", "database", "pd4-RBF.json")).read()),
    }
    path = os.path.join(HERE, "pd4.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB;", len(rows), "structures,", int(out["energy_in"].sum()), "energies,",
          len(out["force_in"]), "force centres")


if __name__ == "__main__":
    main()
