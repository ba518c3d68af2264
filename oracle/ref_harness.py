"""CPU ORACLE (test infrastructure only) — run the UNMODIFIED reference Python in the build container.

``/root/reference`` is a Python package that cannot be pip-installed here (its dependencies
``ase``, ``mpi4py``, ``pyxtal``, ``matplotlib`` are absent and there is no network) and does not
exist on the GPU box.  This harness imports it *in place* (``/root/reference`` on ``sys.path``,
implicit namespace package) after seeding ``sys.modules`` with the minimal stand-ins listed in
SURVEY.md §8(c):

* ``mpi4py.MPI``          size-1 communicator (identity collectives)
* ``ase``, ``ase.db``, ``ase.neighborlist``, ``ase.constraints``, ``ase.calculators.calculator``
                          just enough surface for imports; ``NeighborList`` is a brute-force
                          restatement of ASE's semantics (radii rcut/2, skin 0, bothways,
                          no self interaction, strict ``<``)
* ``pyxtal.database.element.Element``   symbol -> Z
* ``matplotlib``          inert
* ``scipy.special.sph_harm``            shim onto ``sph_harm_y`` (removed in scipy 1.15+)
* ``gpr_calc.kernels._rbf_kernel`` / ``._dot_kernel``   ``lib`` objects backed by ctypes on
                          ``oracle/_ref/*.so`` (the reference C++ compiled by oracle/Makefile), accepting
                          the cffi cdata the reference wrappers pass

It is used ONLY by ``tests/golden/gen_golden.py`` to produce committed golden vectors and by
container-only tests; nothing that runs on the GPU box imports it.
"""
import ctypes
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("GPR_REFERENCE_ROOT", "/root/reference")

SYMBOLS = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As "
           "Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd "
           "Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn").split()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "gpr_calc"))


# ------------------------------------------------------------------------------------------------
# minimal Atoms stand-in (what the reference touches: SURVEY.md Appendix B item 5)
# ------------------------------------------------------------------------------------------------
class _Cell(np.ndarray):
    @property
    def array(self):
        return np.asarray(self)


class _AtomView:
    def __init__(self, number):
        self.number = int(number)


class FixAtoms:
    def __init__(self, indices):
        self.index = np.asarray(indices, dtype=int)

    def get_indices(self):
        return self.index


class Atoms:
    def __init__(self, numbers, positions, cell, pbc=(True, True, True), constraints=None):
        self.numbers = np.asarray(numbers, dtype=int)
        self.positions = np.asarray(positions, dtype=float).copy()
        self.cell = np.asarray(cell, dtype=float).reshape(3, 3).copy().view(_Cell)
        self.pbc = np.asarray(pbc, dtype=bool)
        self.constraints = list(constraints or [])
        self.calc = None

    @property
    def symbols(self):
        return [SYMBOLS[z] for z in self.numbers]

    def __len__(self):
        return len(self.numbers)

    def __getitem__(self, i):
        return _AtomView(self.numbers[i])

    def get_cell(self):
        return self.cell

    def get_volume(self):
        return abs(float(np.linalg.det(np.asarray(self.cell))))

    def get_scaled_positions(self):
        return np.linalg.solve(np.asarray(self.cell).T, self.positions.T).T

    def set_constraint(self, c=None):
        self.constraints = [] if c is None else [c]

    def copy(self):
        return Atoms(self.numbers, self.positions, self.cell, self.pbc, self.constraints)


def neighbor_pairs(positions, cell, pbc, rcut):
    """All (i, j, S) with |r_j + S.cell - r_i| < rcut, excluding (i, i, 0).

    Restates ase.neighborlist.NeighborList(cutoffs=rcut/2, skin=0, bothways=True,
    self_interaction=False) as used at gpr_calc/SO3.py:357-378.  Returned sorted by (i, j, S).
    """
    positions = np.asarray(positions, float)
    cell = np.asarray(cell, float)
    n = len(positions)
    # number of images needed along each periodic axis: rcut / (distance between cell faces)
    nimg = []
    vol = abs(np.linalg.det(cell))
    for k in range(3):
        if pbc[k] and vol > 0:
            a, b = cell[(k + 1) % 3], cell[(k + 2) % 3]
            height = vol / np.linalg.norm(np.cross(a, b))
            nimg.append(int(np.ceil(rcut / height)) + 1)
        else:
            nimg.append(0)
    out = []
    rng = [range(-m, m + 1) for m in nimg]
    for sx in rng[0]:
        for sy in rng[1]:
            for sz in rng[2]:
                shift = np.array([sx, sy, sz]) @ cell
                dvec = positions[None, :, :] + shift[None, None, :] - positions[:, None, :]
                dist = np.sqrt((dvec ** 2).sum(-1))
                ii, jj = np.nonzero(dist < rcut)
                for i, j in zip(ii, jj):
                    if i == j and sx == 0 and sy == 0 and sz == 0:
                        continue
                    out.append((int(i), int(j), sx, sy, sz))
    out.sort()
    return out


class NeighborList:
    def __init__(self, cutoffs, self_interaction=False, bothways=True, skin=0.0):
        assert bothways and not self_interaction and skin == 0.0
        self.rcut = 2 * cutoffs[0]

    def update(self, atoms):
        prs = neighbor_pairs(atoms.positions, np.asarray(atoms.cell), atoms.pbc, self.rcut)
        self._nb = [[] for _ in range(len(atoms))]
        for (i, j, sx, sy, sz) in prs:
            self._nb[i].append((j, (sx, sy, sz)))

    def get_neighbors(self, i):
        nb = self._nb[i]
        if not nb:
            return np.zeros(0, dtype=int), np.zeros((0, 3), dtype=int)
        return np.array([j for j, _ in nb], dtype=int), np.array([s for _, s in nb], dtype=int)


# ------------------------------------------------------------------------------------------------
# stubs
# ------------------------------------------------------------------------------------------------
class _Comm:
    rank = 0

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def bcast(self, x, root=0):
        return x

    def gather(self, x, root=0):
        return [x]

    def scatter(self, x, root=0):
        return x[0]

    def allreduce(self, x, op=None):
        return x

    def barrier(self):
        pass

    Barrier = barrier

    def Reduce(self, send, recv, op=None, root=0):
        if recv is not None:
            np.copyto(recv, send)

    def Allreduce(self, send, recv, op=None):
        s = send[0] if isinstance(send, (list, tuple)) else send
        r = recv[0] if isinstance(recv, (list, tuple)) else recv
        np.copyto(r, s)


class _RefLib:
    """`lib` stand-in for the cffi extension modules: forwards to ctypes, converting cffi cdata."""

    def __init__(self, cdll, sigs):
        import cffi
        self._ffi = cffi.FFI()
        self._cdll = cdll
        self._sigs = sigs

    def __getattr__(self, name):
        sig = self._sigs[name]
        fn = getattr(self._cdll, name)
        ffi = self._ffi

        def call(*args):
            assert len(args) == len(sig), (name, len(args), len(sig))
            conv = []
            for kind, a in zip(sig, args):
                if kind == "i":
                    conv.append(ctypes.c_int(int(a)))
                elif kind == "d":
                    conv.append(ctypes.c_double(float(a)))
                else:
                    conv.append(ctypes.c_void_p(int(ffi.cast("uintptr_t", a))))
            fn(*conv)
        return call


_RBF_SIGS = {
    "rbf_kee_many": "iiiiddd" + "p" * 7,
    "rbf_kee_many_with_grad": "iiiiddd" + "p" * 8,
    "rbf_kef_many": "iiiiddd" + "p" * 8,
    "rbf_kef_many_with_grad": "iiiiddd" + "p" * 8,
    "rbf_kef_many_stress": "iiiiddd" + "p" * 8,
    "rbf_kff_many": "iiiiiidddd" + "p" * 9,
    "rbf_kff_many_with_grad": "iiiiiiddd" + "p" * 10,
    "rbf_kff_many_stress": "iiiiiidddd" + "p" * 9,
}
_DOT_SIGS = {
    "dot_kee_many": "iiiiddd" + "p" * 7,
    "dot_kef_many": "iiiid" + "p" * 8,
    "dot_kef_many_stress": "iiiid" + "p" * 8,
    "dot_kff_many": "iiiiiid" + "p" * 9,
    "dot_kff_many_stress": "iiiiiid" + "p" * 9,
}

_INSTALLED = False


def install():
    """Seed sys.modules and make `import gpr_calc...` resolve to the reference tree."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    here = os.path.dirname(os.path.abspath(__file__))
    from . import kernels as _k
    if not _k.have_ref():
        _k.build(ref=True, port=False)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    MPI = types.SimpleNamespace(COMM_WORLD=_Comm(), SUM="sum", DOUBLE="double")
    mod("mpi4py", MPI=MPI)
    mod("mpi4py.MPI", **MPI.__dict__)

    class Calculator:
        def __init__(self, **kwargs):
            self.parameters = types.SimpleNamespace(**kwargs)
            self.results = {}

        def calculate(self, atoms=None, properties=None, system_changes=None):
            self.atoms = atoms

    mod("ase", Atoms=Atoms)
    mod("ase.db", connect=None)
    mod("ase.neighborlist", NeighborList=NeighborList, PrimitiveNeighborList=NeighborList)
    mod("ase.constraints", FixAtoms=FixAtoms, full_3x3_to_voigt_6_stress=None)
    mod("ase.calculators")
    mod("ase.calculators.calculator", Calculator=Calculator, all_changes=[])

    class Element:
        def __init__(self, sym):
            self.z = SYMBOLS.index(sym)

    mod("pyxtal", pyxtal=None)
    mod("pyxtal.database")
    mod("pyxtal.database.element", Element=Element)
    mpl = mod("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = mod("matplotlib.pyplot")

    import scipy.special as sp
    if not hasattr(sp, "sph_harm"):
        sp.sph_harm = lambda m, l, phi, theta: sp.sph_harm_y(l, m, theta, phi)

    rbf = ctypes.CDLL(os.path.join(here, "_ref", "librbf_ref.so"))
    dot = ctypes.CDLL(os.path.join(here, "_ref", "libdot_ref.so"))
    # `gpr_calc` is seeded explicitly as a namespace rooted at the reference tree: the repository ships an import-path
    # alias package of the same name (gpr_calc/__init__.py -> gpr_calculator_b200), and a regular package anywhere on
    # sys.path would win over the reference's namespace package
    for name in [n for n in sys.modules if n == "gpr_calc" or n.startswith("gpr_calc.")]:
        del sys.modules[name]
    pkg = types.ModuleType("gpr_calc")
    pkg.__path__ = [os.path.join(REF_ROOT, "gpr_calc")]
    sys.modules["gpr_calc"] = pkg
    import gpr_calc.kernels  # noqa: F401
    mod("gpr_calc.kernels._rbf_kernel", lib=_RefLib(rbf, _RBF_SIGS))
    mod("gpr_calc.kernels._dot_kernel", lib=_RefLib(dot, _DOT_SIGS))
    _INSTALLED = True


def modules():
    """Return the reference's own modules (imported under the stubs)."""
    install()
    from gpr_calc.kernels import RBF_mb, Dot_mb, rbf_kernel, dot_kernel, base  # noqa
    from gpr_calc import SO3, utilities  # noqa
    import gpr_calc.gaussianprocess as gp
    return types.SimpleNamespace(RBF_mb=RBF_mb.RBF_mb, Dot_mb=Dot_mb.Dot_mb, rbf_kernel=rbf_kernel,
                                 dot_kernel=dot_kernel, base=base, SO3=SO3.SO3, so3_module=SO3,
                                 utilities=utilities, GP=gp.GP, gp_module=gp)
