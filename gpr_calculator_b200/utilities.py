"""Host-side helpers on the hot path (mirror of the ★ functions of gpr_calc/utilities.py).

Only the pieces the covariance path needs are here: the ragged <-> packed layout
(utilities.py:340-406), descriptor de-duplication (new_pt, :32-42), the conversion of a labelled
structure into training rows (convert_train_data, :97-129), error metrics (:44-63, 81-85) and a
symbol -> atomic-number table (stands in for pyxtal.database.element.Element, used at
gaussianprocess.py:788,847).  Plotting, ASE-db and VASP helpers are out of scope.
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
