"""SO(3) power-spectrum descriptor on the B200 (drop-in for gpr_calc/SO3.py:7-727).

Same constructor arguments, validation (ValueError from the setters, SO3.py:71-145), ``save_dict`` /
``load_from_dict`` and ``calculate(atoms, atom_ids=None, use_mpi=False)`` return dict
``{'x', 'dxdr', 'rdxdr', 'elements', 'seq'}``.  The neighbour search, the expansion coefficients
c_nlm / grad c_nlm and the power spectrum with its derivative are computed by libgpr_b200.so
(csrc/so3.cu); the host only prepares the O(nmax * NQ) radial table.

``calculate_batch`` processes many structures in one pass and can leave the results on device
(what the batched prediction benchmark uses).  Only the cosine cut-off exists in the reference
(the other names at SO3.py:157-170 are undefined there).  stress=True also returns
``rdxdr[n_seq, d, 3, 3] = -pstress / volume`` (SO3.py:253-273, 304-306).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .device import require_cuda, stream, ptr, c_vp


def _radial_overlap_W(nmax):
    """S^(-1/2) of the radial basis overlap matrix (SO3.py:417-430)."""
    a = np.arange(1, nmax + 1)
    t = (2 * a + 5) * (2 * a + 6) * (2 * a + 7)
    ab = a[:, None] + a[None, :]
    S = np.sqrt(t[:, None] * t[None, :]) / ((5 + ab) * (6 + ab) * (7 + ab))
    vals, vecs = np.linalg.eig(np.linalg.inv(S))
    return (vecs @ np.diag(np.sqrt(vals)) @ np.linalg.inv(vecs)).real


def _radial_table(nmax, lmax, rcut, alpha):
    """Gauss-Chebyshev nodes rho_q on (0, rcut) and G[n, q] = g_n(rho) rho^2 e^{-a rho^2} sqrt(1-t^2) w
    (SO3.py:432-453, 619-633, 646-647)."""
    NQ = (nmax + lmax + 1) * 10
    q = np.arange(1, NQ + 1)
    t = np.cos((2 * q - 1) * np.pi / 2 / NQ)
    w = np.pi / NQ * rcut / 2
    rho = rcut / 2 * (t + 1)
    a = np.arange(1, nmax + 1)[:, None]
    phi = (rcut - rho[None, :]) ** (a + 2) / np.sqrt(2 * rcut ** (2 * a + 7) / (2 * a + 5) / (2 * a + 6) / (2 * a + 7))
    g = _radial_overlap_W(nmax) @ phi
    return rho, g * w * rho ** 2 * np.exp(-alpha * rho ** 2) * np.sqrt(1 - t ** 2)


class SO3:
    '''
    SO(3) power spectrum of the Gaussian-smoothed neighbour density ("On Representing Atomic
    Environments"), with Cartesian derivatives.

    args:
        nmax: int, degree of radial expansion
        lmax: int, degree of spherical harmonic expansion
        rcut: float, cutoff radius for neighbor calculation
        alpha: float, gaussian width parameter
        derivative: bool, whether to calculate the gradient of not
        weight_on: bool, if True, neighbours of a different species count negatively
        primitive: bool, kept for signature compatibility (the device search has one code path)
    '''

    def __init__(self, nmax=3, lmax=3, rcut=3.5, alpha=2.0, derivative=True, stress=False,
                 cutoff_function='cosine', weight_on=False, primitive=False):
        self.nmax = nmax
        self.lmax = lmax
        self.rcut = rcut
        self.alpha = alpha
        self.derivative = derivative
        self.stress = stress
        self._type = "SO3"
        self.cutoff_function = cutoff_function
        self.weight_on = weight_on
        self.primitive = primitive
        self._tables = None

    def __str__(self):
        s = "SO3 descriptor with Cutoff: {:6.3f}".format(self.rcut)
        s += " lmax: {:d}, nmax: {:d}, alpha: {:.3f}\n".format(self.lmax, self.nmax, self.alpha)
        return s

    def __repr__(self):
        return str(self)

    def load_from_dict(self, dict0):
        self.nmax = dict0["nmax"]
        self.lmax = dict0["lmax"]
        self.rcut = dict0["rcut"]
        self.alpha = dict0["alpha"]
        self.derivative = dict0["derivative"]
        self.stress = dict0["stress"]

    def save_dict(self):
        return {"nmax": self.nmax, "lmax": self.lmax, "rcut": self.rcut, "alpha": self.alpha,
                "derivative": self.derivative, "stress": self.stress, "_type": "SO3"}

    # -- validated attributes (same messages as SO3.py:71-145) -----------------------------------
    @property
    def nmax(self):
        return self._nmax

    @nmax.setter
    def nmax(self, nmax):
        if isinstance(nmax, int):
            if nmax < 1:
                raise ValueError('nmax must be greater than or equal to 1')
            if nmax > 11:
                raise ValueError('nmax > 11 yields complex eigenvalues which will mess up the calculation')
            self._nmax = nmax
            self._tables = None
        else:
            raise ValueError('nmax must be an integer')

    @property
    def lmax(self):
        return self._lmax

    @lmax.setter
    def lmax(self, lmax):
        if isinstance(lmax, int):
            if lmax < 0:
                raise ValueError('lmax must be greater than or equal to zero')
            elif lmax > 15:       # the reference stops at 32 (SO3.py:127); the device kernels hold Y_lm tables up to l = 16
                raise NotImplementedError('lmax > 15 is not supported by the device kernels')
            self._lmax = lmax
            self._tables = None
        else:
            raise ValueError('lmax must be an integer')

    @property
    def rcut(self):
        return self._rcut

    @rcut.setter
    def rcut(self, rcut):
        if isinstance(rcut, (float, int)):
            if rcut <= 0:
                raise ValueError('rcut must be greater than zero')
            self._rcut = rcut
            self._tables = None
        else:
            raise ValueError('rcut must be a float')

    @property
    def alpha(self):
        return self._alpha

    @alpha.setter
    def alpha(self, alpha):
        if isinstance(alpha, (float, int)):
            if alpha <= 0:
                raise ValueError('alpha must be greater than zero')
            self._alpha = alpha
            self._tables = None
        else:
            raise ValueError('alpha must be a float')

    @property
    def derivative(self):
        return self._derivative

    @derivative.setter
    def derivative(self, derivative):
        if isinstance(derivative, bool):
            self._derivative = derivative
        else:
            raise ValueError('derivative must be a boolean value')

    @property
    def stress(self):
        return self._stress

    @stress.setter
    def stress(self, stress):
        if isinstance(stress, bool):
            self._stress = stress
        else:
            raise ValueError('stress must be a boolean value')

    @property
    def cutoff_function(self):
        return self._cutoff_function

    @cutoff_function.setter
    def cutoff_function(self, cutoff_function):
        if isinstance(cutoff_function, str):
            if cutoff_function != 'cosine':
                raise NotImplementedError('The requested cutoff function has not been implemented')
            self._cutoff_function = 'cosine'
        else:
            raise ValueError('You must specify the cutoff function as a string')

    # ---------------------------------------------------------------------------------------------
    @property
    def ncoefs(self):
        return self.nmax * (self.nmax + 1) // 2 * (self.lmax + 1)

    def _device_tables(self):
        if self._tables is None:
            rho, G = _radial_table(self.nmax, self.lmax, float(self.rcut), float(self.alpha))
            ls = np.arange(self.lmax + 1)
            norm = np.sqrt(2 * np.sqrt(2) * np.pi / np.sqrt(2 * ls + 1))        # SO3.py:206
            self._tables = tuple(torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
                                 for a in (rho, G, norm))
        return self._tables

    @staticmethod
    def _images(cell, pbc, rcut):
        """Images to search along each axis: ceil(rcut / height) + 1 on periodic axes."""
        cell = np.asarray(cell, float).reshape(3, 3)
        vol = abs(np.linalg.det(cell))
        out = []
        for k in range(3):
            if pbc[k] and vol > 0:
                a, b = cell[(k + 1) % 3], cell[(k + 2) % 3]
                out.append(int(np.ceil(rcut / (vol / np.linalg.norm(np.cross(a, b))))) + 1)
            else:
                out.append(0)
        return out

    def calculate_batch(self, structures, to_host=True):
        """Descriptors of a list of structures in one device pass.

        Returns a list of dicts (one per structure) when to_host, else a dict of device tensors
        ``{'x' [A,d], 'dxdr' [Q,d,3], 'seq' [Q,2] (local atom ids), 'atom_ptr', 'seq_ptr', 'numbers'}``.
        """
        require_cuda()
        if self.stress and not self.derivative:
            raise ValueError("stress=True needs derivative=True")
        if self.lmax > 15:
            raise NotImplementedError("the device kernels support lmax <= 15")
        S = len(structures)
        counts = [len(s) for s in structures]
        atom_ptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int32)
        A = int(atom_ptr[-1])
        pos = np.concatenate([np.asarray(s.positions, dtype=np.float64).reshape(-1, 3) for s in structures]) if A else np.zeros((0, 3))
        cells = np.stack([np.asarray(s.cell, dtype=np.float64).reshape(3, 3) for s in structures]).reshape(S, 9)
        nimg = np.array([self._images(np.asarray(s.cell), np.asarray(s.pbc, dtype=bool), float(self.rcut)) for s in structures],
                        dtype=np.int32).reshape(S, 3)
        numbers = np.concatenate([np.asarray(s.numbers, dtype=np.int32) for s in structures]) if A else np.zeros(0, np.int32)
        struct_of = np.repeat(np.arange(S, dtype=np.int32), counts)

        dev = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")  # noqa: E731
        t_atom_ptr, t_struct_of = dev(atom_ptr, torch.int32), dev(struct_of, torch.int32)
        t_pos, t_cell, t_nimg, t_num = dev(pos, torch.float64), dev(cells, torch.float64), dev(nimg, torch.int32), dev(numbers, torch.int32)
        rho, G, norm = self._device_tables()
        st = stream()
        nnb = torch.zeros(A, dtype=torch.int32, device="cuda")
        nuniq = torch.zeros(A, dtype=torch.int32, device="cuda")
        _lib.call("gprb_so3_neighbors", S, A, ptr(t_atom_ptr), ptr(t_struct_of), ptr(t_pos), ptr(t_cell), ptr(t_nimg),
                  float(self.rcut), 0, ptr(nnb), ptr(nuniq), c_vp(0), c_vp(0), c_vp(0), st)
        nb_ptr = torch.zeros(A + 1, dtype=torch.int32, device="cuda")
        seq_ptr = torch.zeros(A + 1, dtype=torch.int32, device="cuda")
        nb_ptr[1:] = torch.cumsum(nnb, 0)
        seq_ptr[1:] = torch.cumsum(nuniq, 0)
        n_nb, n_seq = (int(v) for v in torch.stack((nb_ptr[-1], seq_ptr[-1])).cpu())
        nb_j = torch.empty(max(n_nb, 1), dtype=torch.int32, device="cuda")
        nb_rvec = torch.empty((max(n_nb, 1), 3), dtype=torch.float64, device="cuda")
        _lib.call("gprb_so3_neighbors", S, A, ptr(t_atom_ptr), ptr(t_struct_of), ptr(t_pos), ptr(t_cell), ptr(t_nimg),
                  float(self.rcut), 1, c_vp(0), c_vp(0), ptr(nb_ptr), ptr(nb_j), ptr(nb_rvec), st)
        nnl = self.nmax * (self.lmax + 1)
        rad = torch.empty((max(n_nb, 1), 2 * nnl), dtype=torch.float64, device="cuda")
        _lib.call("gprb_so3_radial", n_nb, ptr(nb_rvec), self.nmax, self.lmax, int(rho.numel()), float(self.alpha),
                  float(self.rcut), ptr(rho), ptr(G), ptr(rad), st)
        d = self.ncoefs
        x = torch.zeros((A, d), dtype=torch.float64, device="cuda")
        dxdr = torch.zeros((n_seq, d, 3), dtype=torch.float64, device="cuda") if self.derivative else None
        seq = torch.zeros((n_seq, 2), dtype=torch.int64, device="cuda") if self.derivative else None
        rdxdr = inv_vol = None
        if self.stress:
            # rdxdr = -pstress / volume (SO3.py:304-306)
            rdxdr = torch.zeros((n_seq, d, 3, 3), dtype=torch.float64, device="cuda")
            inv_vol = dev(1.0 / np.abs(np.linalg.det(cells.reshape(S, 3, 3))), torch.float64)
        _lib.call("gprb_so3_power", A, ptr(nb_ptr), ptr(nb_j), ptr(nb_rvec), ptr(rad), ptr(t_num), ptr(t_atom_ptr),
                  ptr(t_struct_of), ptr(seq_ptr), self.nmax, self.lmax, float(self.alpha), float(self.rcut), ptr(norm),
                  (1 if self.derivative else 0) | (2 if self.weight_on else 0), ptr(x), ptr(dxdr), ptr(seq), ptr(t_pos),
                  ptr(inv_vol), ptr(rdxdr), st)
        if not to_host:
            return {'x': x, 'dxdr': dxdr, 'rdxdr': rdxdr, 'seq': seq, 'atom_ptr': t_atom_ptr, 'seq_ptr': seq_ptr,
                    'numbers': t_num, 'n_neighbors': n_nb}
        xh = x.cpu().numpy()
        out = []
        if self.derivative:
            dh, sh = dxdr.cpu().numpy(), seq.cpu().numpy()
            seq_atom = seq_ptr.cpu().numpy()
            rh_ = rdxdr.cpu().numpy() if rdxdr is not None else None
        for k, s in enumerate(structures):
            a0, a1 = int(atom_ptr[k]), int(atom_ptr[k + 1])
            item = {'x': xh[a0:a1], 'dxdr': None, 'rdxdr': None, 'elements': list(s.symbols)}
            if self.derivative:
                q0, q1 = int(seq_atom[a0]), int(seq_atom[a1])
                item['dxdr'] = dh[q0:q1]
                item['seq'] = sh[q0:q1]
                if rh_ is not None:
                    item['rdxdr'] = rh_[q0:q1]
            out.append(item)
        return out

    def calculate(self, atoms, atom_ids=None, use_mpi=False):
        '''
        Power spectrum components (and derivatives) of one structure (SO3.py:186-323).

        Args:
            atoms: an ASE-like atoms object (positions, cell, pbc, numbers, symbols)
            atom_ids: centre atoms to evaluate (None = all).  As in build_neighbor_list (SO3.py:354-401) the rows of
                      the other atoms stay zero in x, and seq / dxdr / rdxdr hold the (i, j) rows of the requested
                      centres only, in the order the ids are given.
            use_mpi: ignored; the device pass replaces the MPI split of SO3.py:228-296
        '''
        out = self.calculate_batch([atoms], to_host=True)[0]
        if atom_ids is None:
            return out
        ids = [int(i) for i in atom_ids]
        x = np.zeros_like(out['x'])
        x[ids] = out['x'][ids]
        out['x'] = x
        if self.derivative:
            seq = out['seq']
            rows = np.concatenate([np.flatnonzero(seq[:, 0] == i) for i in ids]) if ids else np.zeros(0, dtype=np.int64)
            out['seq'] = seq[rows]
            out['dxdr'] = out['dxdr'][rows]
            if out.get('rdxdr') is not None:
                out['rdxdr'] = out['rdxdr'][rows]
        return out
