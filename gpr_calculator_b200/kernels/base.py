"""Host-side block assembly (mirror of gpr_calc/kernels/base.py:3-38)."""
import numpy as np


def build_covariance(c_ee, c_ef, c_fe, c_ff, c_se=None, c_sf=None):
    """Stack whichever of the four blocks exist, with the reference's dispatch table
    (kernels/base.py:3-30): full 2x2, one block row, or a single block."""
    have = tuple(x is not None for x in (c_ee, c_ef, c_fe, c_ff))
    if all(have):
        return np.block([[c_ee, c_ef], [c_fe, c_ff]])
    table = {
        (False, False, True, True): lambda: np.hstack((c_fe, c_ff)),
        (True, True, False, False): lambda: np.hstack((c_ee, c_ef)),
        (False, True, False, False): lambda: c_ef,
        (True, False, False, False): lambda: c_ee,
        (False, False, False, True): lambda: c_ff,
        (False, False, True, False): lambda: c_fe,
    }
    fn = table.get(have)
    return fn() if fn is not None else None


def get_mask(ele1, ele2):
    """Index pairs whose species differ, or None (kernels/base.py:32-38)."""
    ids = np.where((ele1[:, None] - ele2[None, :]) != 0)
    return None if len(ids[0]) == 0 else ids
