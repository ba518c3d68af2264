"""Row-block sharding of the covariance matrix over GPUs (one process per GPU, NCCL via torch.distributed).

Replaces the reference's mpi4py pattern — every rank computes a row slab, rank 0 gathers pickles,
vstacks and broadcasts (RBF_mb.py:471-521, gaussianprocess.py:246-247,305-306) — with:
  * contiguous row blocks of training groups balanced by cost (rows of the group),
  * one in-place all-gather of the slabs into the full K on every GPU,
  * dK/dtheta never gathered: each rank traces its own rows, scalars are all-reduced.
With the gloo backend (CPU tensors) the same partition / gather logic is exercised in the tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def split_groups(costs, parts, speed=None):
    """Contiguous partition of len(costs) groups into `parts` blocks with balanced total cost
    (block p gets a share proportional to speed[p] when given).

    Returns `parts + 1` boundaries b with b[0] = 0 and b[-1] = len(costs)."""
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if n == 0:
        return [0] * (parts + 1)
    cum = np.concatenate(([0.0], np.cumsum(costs)))
    total = cum[-1]
    share = np.ones(parts) if speed is None else np.asarray(speed, dtype=np.float64)
    share = np.cumsum(share) / share.sum()
    bounds = [0]
    for p in range(1, parts):
        target = total * share[p - 1]
        j = int(np.searchsorted(cum, target, side="left"))
        if j > 0 and abs(cum[j - 1] - target) <= abs(cum[min(j, n)] - target):
            j -= 1
        j = min(max(j, bounds[-1]), n)
        bounds.append(j)
    bounds.append(n)
    return bounds


def row_windows(e_rows, f_rows, parts, upper=False, speed=None):
    """Per-rank windows ((e0,e1),(f0,f1)) over energy groups and force groups.

    Energy and force groups are partitioned independently so that every rank gets an equal share
    of both block rows (the force block dominates: cost ~ rows of the centre).  upper=True balances
    the trapezoids of an upper-triangle build (GPRB_FF_UPPER): force group I costs
    n_I * sum_{J >= I} n_J.  speed[r] (relative measured throughput of rank r, see GP._build_K) scales
    the share of the force rows rank r receives."""
    eb = split_groups(e_rows, parts)
    f = np.asarray(f_rows, dtype=np.float64)
    if upper and len(f):
        f = f * np.cumsum(f[::-1])[::-1]
    fb = split_groups(f, parts, speed=speed)
    return [((eb[r], eb[r + 1]), (fb[r], fb[r + 1])) for r in range(parts)]


def window_row_ranges(windows, NE):
    """Row ranges in the assembled matrix covered by each window: list of [(lo,hi) energy, (lo,hi) force]."""
    return [((e0, e1), (NE + 3 * f0, NE + 3 * f1)) for (e0, e1), (f0, f1) in windows]


def gather_rows(K_local, windows, NE, N, group=None):
    """All-gather row slabs into the full [N, N_cols] matrix on every rank.

    K_local holds this rank's rows: first its energy rows, then its force rows."""
    rank, size = world()
    n_cols = K_local.shape[1]
    if size == 1:
        return K_local
    K = torch.empty((N, n_cols), dtype=K_local.dtype, device=K_local.device)
    (e0, e1), (f0, f1) = windows[rank]
    K[e0:e1] = K_local[:e1 - e0]
    K[NE + 3 * f0:NE + 3 * f1] = K_local[e1 - e0:]
    gather_rows_inplace(K, windows, NE, group=group)
    return K


def gather_rows_inplace(K, windows, NE, group=None):
    """In-place all-gather: every rank has written its own energy and force row slabs of the full
    matrix K; on return every rank holds all rows.  The slabs are contiguous row ranges of the
    row-major matrix, so each is sent from and received into K itself (no staging copies); uneven
    slab sizes are handled by the backend (NCCL: grouped broadcasts)."""
    rank, size = world()
    if size == 1:
        return K
    for slabs in ([K[e0:e1] for (e0, e1), _ in windows], [K[NE + 3 * f0:NE + 3 * f1] for _, (f0, f1) in windows]):
        if all(s.numel() == 0 for s in slabs):
            continue
        n0 = slabs[0].numel()
        lo = slabs[0].data_ptr()
        if n0 > 0 and all(s.numel() == n0 and s.data_ptr() == lo + r * n0 * K.element_size() for r, s in enumerate(slabs)):
            # equal slabs tiling one contiguous range: a single all-gather straight into K
            whole = K.view(-1)[(lo - K.data_ptr()) // K.element_size():][:size * n0]
            dist.all_gather_into_tensor(whole, slabs[rank].reshape(-1).clone(), group=group)
            continue
        # uneven slabs (cost-balanced windows): one broadcast per owner, each received in place
        for r, s in enumerate(slabs):
            if s.numel():
                dist.broadcast(s, src=r if group is None else dist.get_global_rank(group, r), group=group)
    return K


def all_gather_floats(value, device="cpu", group=None):
    """One python float per rank -> list of all ranks' values."""
    rank, size = world()
    if size == 1:
        return [float(value)]
    t = torch.zeros(size, dtype=torch.float64, device=device)
    t[rank] = float(value)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(v) for v in t.cpu()]


def all_reduce_sum(values, device="cpu", group=None):
    """Sum a small list of python floats over ranks."""
    rank, size = world()
    if size == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(v) for v in t.cpu()]
