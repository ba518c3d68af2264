"""CPU, world_size 2, gloo: the row-block gather / scalar all-reduce logic of the multi-GPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, size, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    from gpr_calculator_b200 import dist as gd
    NE, NF = 5, 11
    N = NE + 3 * NF
    full = torch.arange(N * N, dtype=torch.float64).reshape(N, N)
    windows = gd.row_windows([10] * NE, list(range(20, 20 + NF)), size)
    (e0, e1), (f0, f1) = windows[rank]
    local = torch.cat((full[e0:e1], full[NE + 3 * f0:NE + 3 * f1]))
    got = gd.gather_rows(local, windows, NE, N)
    ok = bool(torch.equal(got, full))
    # in-place variant: every rank has written only its own slabs of the full matrix
    mine = torch.full((N, N), -1.0, dtype=torch.float64)
    mine[e0:e1] = full[e0:e1]
    mine[NE + 3 * f0:NE + 3 * f1] = full[NE + 3 * f0:NE + 3 * f1]
    gd.gather_rows_inplace(mine, windows, NE)
    ok = ok and bool(torch.equal(mine, full))
    wu = gd.row_windows([10] * NE, list(range(20, 20 + NF)), size, upper=True)
    ok = ok and wu[0][1][0] == 0 and wu[-1][1][1] == NF and wu[0][1][1] <= windows[0][1][1]   # early rows cost more
    # sharded prediction: every rank holds the results of its contiguous block of structures
    n_atoms = [3, 1, 4, 1, 5, 9, 2]
    truth = [(float(k), np.full((n, 3), k + 0.5), None, 10.0 + k, np.full((n, 3), k + 0.25)) for k, n in enumerate(n_atoms)]
    b = gd.split_groups(n_atoms, size)
    for std in (False, True):
        got_p = gd.gather_predictions([t if std else t[:3] for t in truth[b[rank]:b[rank + 1]]], n_atoms, b[rank], std)
        for t, g in zip(truth, got_p):
            ok = ok and g[0] == t[0] and np.array_equal(g[1], t[1]) and g[2] is None and len(g) == (5 if std else 3)
            if std:
                ok = ok and g[3] == t[3] and np.array_equal(g[4], t[4])
    ok = ok and gd.broadcast_floats([float(rank), 7.0], src=0) == [0.0, 7.0]
    # re-cut of row slabs between two contiguous partitions (dK/dl rows for the inverse-rows gradient)
    rows_all = torch.arange(13 * 4, dtype=torch.float64).reshape(13, 4)
    old_b, new_b = [0, 9, 13], [0, 4, 13]
    send = [max(0, min(old_b[rank + 1], new_b[d + 1]) - max(old_b[rank], new_b[d])) for d in range(size)]
    recv = [max(0, min(old_b[s_ + 1], new_b[rank + 1]) - max(old_b[s_], new_b[rank])) for s_ in range(size)]
    got_r = gd.all_to_all_rows(torch.empty((new_b[rank + 1] - new_b[rank], 4), dtype=torch.float64),
                               rows_all[old_b[rank]:old_b[rank + 1]].clone(), recv, send)
    ok = ok and bool(torch.equal(got_r, rows_all[new_b[rank]:new_b[rank + 1]]))
    # shared Cholesky: ownership of the block columns, panel broadcasts and trailing updates with numpy stand-ins for the two
    # library steps (gprb_chol_panel / gprb_chol_trailing); every rank must end with the full factor in its lower triangle
    rng = np.random.default_rng(3)
    A = rng.normal(size=(53, 53))
    Kn = A @ A.T + 53 * np.eye(53)
    Kt = torch.from_numpy(Kn.copy())

    def panel(K, k0, nbk, info):
        a = K.numpy()
        L = np.linalg.cholesky(a[k0:k0 + nbk, k0:k0 + nbk])
        a[k0:k0 + nbk, k0:k0 + nbk] = L
        a[k0 + nbk:, k0:k0 + nbk] = np.linalg.solve(L, a[k0 + nbk:, k0:k0 + nbk].T).T

    def trailing(K, k0, nbk, j0, nbj):
        a = K.numpy()
        a[j0:, j0:j0 + nbj] -= a[j0:, k0:k0 + nbk] @ a[j0:j0 + nbj, k0:k0 + nbk].T

    info = gd.distributed_cholesky(Kt, nb=8, panel_fn=panel, trailing_fn=trailing)
    ok = ok and info == 0 and bool(np.allclose(np.tril(Kt.numpy()), np.linalg.cholesky(Kn), rtol=0, atol=1e-12))
    s = gd.all_reduce_sum([float(rank + 1), 2.0])
    q.put((rank, ok, s, gd.world()))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_rows_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, ok, s, w in res:
        assert ok, "rank %d reassembled a wrong matrix" % rank
        assert s == [3.0, 4.0] and w == (rank, 2)


def test_single_rank_is_identity():
    from gpr_calculator_b200 import dist as gd
    K = torch.ones((4, 4), dtype=torch.float64)
    assert gd.gather_rows(K, [((0, 1), (0, 1))], 1, 4) is K
    assert gd.all_reduce_sum([1.5]) == [1.5] and gd.world() == (0, 1)
