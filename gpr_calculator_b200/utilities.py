"""Host-side helpers on the hot path (mirror of the ★ functions of gpr_calc/utilities.py).

Only the pieces the covariance path needs are here: the ragged <-> packed layout
(utilities.py:340-406), descriptor de-duplication (new_pt, :32-42), the conversion of a labelled
structure into training rows (convert_train_data, :97-129), error metrics (:44-63, 81-85) and a
symbol -> atomic-number table (stands in for pyxtal.database.element.Element, used at
gaussianprocess.py:788,847), and the database readers that produce training rows (get_data, convert_struc,
get_train_data, get_strucs, :132-246).  Plotting and VASP helpers are out of scope.
"""
import numpy as np

_SYMBOLS = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As "
            "Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd "
            "Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am "
            "Cm Bk Cf Es Fm Md No Lr").split()
_Z = {s: i for i, s in enumerate(_SYMBOLS)}


def atomic_number(symbol):
    """Element(symbol).z of pyxtal."""
    try:
        return _Z[symbol]
    except KeyError:
        raise ValueError("unknown chemical symbol %r" % (symbol,))


def atomic_numbers(symbols):
    return np.array([atomic_number(s) for s in symbols], dtype=np.int64)


def chemical_symbol(z):
    return _SYMBOLS[int(z)]


# ------------------------------------------------------------------------------------------------
# ragged <-> packed layout
# ------------------------------------------------------------------------------------------------
def list_to_tuple(data, stress=False, include_value=False, mode='force'):
    """Stack per-group arrays into the packed layout the kernels consume.

    force mode : items (x, dxdr[, f], ele) -> (X, dXdR, ELE, indices[, values])
    energy mode: items (x[, e], ele)       -> (X, ELE, indices[, values])
    Same signature and return convention as gpr_calc/utilities.py:340-390.
    """
    data = list(data)
    if len(data) == 0:
        raise ValueError("list_to_tuple: empty data")
    indices = [int(item[0].shape[0]) for item in data]
    ncoef = int(data[-1][0].shape[1])
    total = sum(indices)
    X = np.zeros([total, ncoef])
    ELE = []
    values = []
    force = mode == 'force'
    if force:
        dXdR = np.zeros([total, ncoef, 9 if stress else 3])
    pos = 0
    for item, n in zip(data, indices):
        if force:
            if include_value:
                x, dxdr, val, ele = item
            else:
                x, dxdr, ele = item
            dXdR[pos:pos + n] = dxdr
        else:
            if include_value:
                x, val, ele = item
            else:
                x, ele = item
        if include_value:
            values.append(val)
        X[pos:pos + n] = x
        ELE.extend(ele)
        pos += n
    ELE = np.ravel(ELE)
    out = (X, dXdR, ELE, indices) if force else (X, ELE, indices)
    return out + (values,) if include_value else out


def tuple_to_list(data, mode='force'):
    """Inverse of list_to_tuple (utilities.py:393-406)."""
    out = []
    start = 0
    if mode == 'force':
        X, dXdR, ELE, indices = data
        for n in indices:
            out.append((X[start:start + n], dXdR[start:start + n], ELE[start:start + n]))
            start += n
    else:
        X, ELE, indices = data
        for n in indices:
            out.append((X[start:start + n], ELE[start:start + n]))
            start += n
    return out


# ------------------------------------------------------------------------------------------------
# selection helpers
# ------------------------------------------------------------------------------------------------
def new_pt(data, Refs, d_tol=1e-1, eps=1e-8):
    """True if descriptor `data=(X, ele)` is not within 1-cos^2 < d_tol of any same-species reference.

    Same arithmetic as utilities.py:32-42 (including where eps enters each normalisation)."""
    X, ele = data
    X = X / (np.linalg.norm(X) + eps)
    for X1, ele1 in Refs:
        if ele1 == ele:
            X1 = X1 / np.linalg.norm(X1 + eps)
            c = X @ X1.T
            if 1 - c ** 2 < d_tol:
                return False
    return True


def force_rows(d, ele, i):
    """Rows of the descriptor dict that make up the force datum of atom i:
    all (centre, i) pairs of `seq` (gaussianprocess.py:857-861, utilities.py:114-117)."""
    ids = np.argwhere(d['seq'][:, 1] == i).flatten()
    centres = d['seq'][ids, 0]
    return d['x'][centres, :], d['dxdr'][ids], ele[centres]


def convert_train_data(data, des, N_force=100000):
    """[(struc, energy, forces), ...] -> {'energy': [...], 'force': [...], 'db': [...]}.

    Mirrors utilities.py:97-129: every atom contributes a force datum (up to N_force in total; the
    reference's de-duplication branch there is unreachable), the energy datum is per atom."""
    energy_data, force_data, db_data = [], [], []
    for struc, energy, forces in data:
        d = des.calculate(struc)
        ele = atomic_numbers(d['elements'])
        f_ids = []
        for i in range(len(struc)):
            if len(force_data) < N_force:
                x, dxdr, e = force_rows(d, ele, i)
                force_data.append((x, dxdr, forces[i], e))
                f_ids.append(i)
        energy_data.append((d['x'], energy / len(struc), ele))
        db_data.append((struc, energy, forces, True, f_ids))
    return {"energy": energy_data, "force": force_data, "db": db_data}


# ------------------------------------------------------------------------------------------------
# database -> training rows (utilities.py:132-246): the input side of the path.  The ASE sqlite file is
# read by asedb.py (no ASE needed); descriptors of all selected structures come from one batched device pass
# when the descriptor offers calculate_batch (the reference loops des.calculate or forks a Pool, :207-224).
# ------------------------------------------------------------------------------------------------
def _rows(db_file):
    from .asedb import read_rows
    return read_rows(db_file)


def get_train_data(db_file, include_stress=False):
    """(structures, energies, forces[, stresses]) of every row (utilities.py:166-182)."""
    strucs, energies, forces, stresses = [], [], [], []
    for row in _rows(db_file):
        strucs.append(row.toatoms())
        energies.append(row.data["energy"])
        forces.append(np.array(row.data["force"]))
        if include_stress:
            stresses.append(np.array(row.data["stress"]))
    return (strucs, energies, forces, stresses) if include_stress else (strucs, energies, forces)


def get_strucs(db_file, N_max=None):
    """structures and (E, F, S or None) per row (utilities.py:225-241)."""
    structures, values = [], []
    for row in _rows(db_file):
        structures.append(row.toatoms())
        S = np.array(row.data["stress"]) if "stress" in row.data else None
        values.append((row.data["energy"], np.array(row.data["force"]), S))
        if N_max is not None and len(values) == N_max:
            break
    return structures, values


def fea(des, struc):
    return des.calculate(struc)


def convert_struc(db_file, des, ids=None, N=None, ncpu=1, stress=False, batch=16):
    """Descriptors and labels of the rows of a database (utilities.py:185-223): rows whose 0-based position is in
    `ids` (all when None), at most N of them.  Returns (list of descriptor dicts, {'energy','forces','stress'}, structures).
    `ncpu` is accepted for signature compatibility: the device pass replaces the process pool."""
    structures, train_Y = [], {"energy": [], "forces": [], "stress": []}
    for row in _rows(db_file):
        if ids is not None and (row.id - 1) not in ids:
            continue
        train_Y["energy"].append(row.data["energy"])
        train_Y["forces"].append(np.array(row.data["force"]))
        if stress:
            train_Y["stress"].append(np.array(row.data["stress"]))
        structures.append(row.toatoms())
        if N is not None and len(structures) == N:
            break
    if hasattr(des, "calculate_batch"):
        xs = []
        for s0 in range(0, len(structures), batch):
            xs += des.calculate_batch(structures[s0:s0 + batch], to_host=True)
    else:
        xs = [des.calculate(struc) for struc in structures]
    return xs, train_Y, structures


def get_data(db_name, des, N_force=100000, lists=None, select=False, no_energy=False, ncpu=1):
    """Training dict {'energy', 'force', 'db'} from a database (utilities.py:132-163): energy item (x, E / n_atoms, Z),
    force item (x[seq[ids, 0]], dxdr[ids], F[i], Z[seq[ids, 0]]) with ids = argwhere(seq[:, 1] == i) for every atom i
    (only atom 0 when select=True), at most N_force force items in total."""
    X, Y, structures = convert_struc(db_name, des, lists, ncpu=ncpu)
    energy_data, force_data, db_data = [], [], []
    for k in range(len(X)):
        ele = atomic_numbers(X[k]['elements'])
        energy_data.append((X[k]['x'], Y["energy"][k] / len(X[k]['x']), ele))
        f_ids = []
        for i in ([0] if select else range(len(X[k]['x']))):
            if len(force_data) < N_force:
                x, dxdr, e = force_rows(X[k], ele, i)
                force_data.append((x, dxdr, Y['forces'][k][i], e))
                f_ids.append(i)
        db_data.append((structures[k], Y['energy'][k], Y['forces'][k], True, f_ids))
    return {"energy": [] if no_energy else energy_data, "force": force_data, "db": db_data}


# ------------------------------------------------------------------------------------------------
# error metrics (utilities.py:44-63, 81-85)
# ------------------------------------------------------------------------------------------------
def rmse(true, predicted):
    true, predicted = np.array(true), np.array(predicted)
    return np.sqrt(sum((true - predicted) ** 2 / len(true)))


def mae(true, predicted):
    true, predicted = np.array(true), np.array(predicted)
    return sum(abs(true - predicted) / len(true))


def r2(true, predicted):
    if len(true) == 0:
        return 1
    true, predicted = np.array(true), np.array(predicted)
    mean = sum(true) / len(true)
    return 1 - sum((true - predicted) ** 2) / (sum((true - mean) ** 2) + 1e-8)


def metrics(y_train, y_test, y_train_pred, y_test_pred, header):
    """Two printed lines 'header Train[ n]: R2 .. MAE .. RMSE ..' (utilities.py:65-79)."""
    out = []
    for tag, y, yp in (("Train", y_train, y_train_pred), ("Test ", y_test, y_test_pred)):
        r2_, mae_, rmse_ = metric_values(y, yp)
        out.append("{:s} {:s}[{:4d}]: R2 {:6.4f} MAE {:6.3f} RMSE {:6.3f}".format(header, tag, len(y), r2_, mae_, rmse_))
        print(out[-1])
    return tuple(out)


def metric_single(y_train, y_train_pred, header, show_max=False):
    """One printed line (utilities.py:87-95; the reference formats the floats with '{:s}' and raises -- the values
    are printed here the way `metrics` prints them)."""
    r2_, mae_, rmse_ = metric_values(y_train, y_train_pred)
    line = "{:s} [{:4d}]: R2 {:6.4f} MAE {:6.3f} RMSE {:6.3f}".format(header, len(y_train), r2_, mae_, rmse_)
    if show_max:
        line += '  Max {:6.4f}'.format(np.max(np.abs(np.asarray(y_train_pred) - np.asarray(y_train))))
    print(line)
    return line


def metric_values(y, y_pred):
    return r2(y, y_pred), mae(y, y_pred), rmse(y, y_pred)


class SimpleAtoms:
    """Minimal Atoms container (positions, cell, pbc, numbers) for synthetic benchmarks and tests.

    Real ase.Atoms objects are accepted everywhere instead; this class only exists because ASE is
    not a dependency of the hot path."""

    def __init__(self, numbers, positions, cell, pbc=(True, True, True), constraints=None):
        self.numbers = np.asarray(numbers, dtype=np.int64)
        self.positions = np.array(positions, dtype=np.float64)
        self.cell = np.array(cell, dtype=np.float64).reshape(3, 3)
        self.pbc = np.asarray(pbc, dtype=bool)
        self.constraints = list(constraints or [])
        self.calc = None

    @property
    def symbols(self):
        return [chemical_symbol(z) for z in self.numbers]

    def __len__(self):
        return len(self.numbers)

    def get_cell(self):
        return self.cell

    def get_volume(self):
        return abs(float(np.linalg.det(self.cell)))

    def copy(self):
        return SimpleAtoms(self.numbers, self.positions, self.cell, self.pbc, self.constraints)

    def get_potential_energy(self):
        return self.calc.get_potential_energy(self)

    def get_forces(self):
        return self.calc.get_forces(self)


class FixAtoms:
    """Stand-in for ase.constraints.FixAtoms (only get_indices is used, gaussianprocess.py:823-832)."""

    def __init__(self, indices):
        self.index = np.asarray(indices, dtype=int)

    def get_indices(self):
        return self.index
