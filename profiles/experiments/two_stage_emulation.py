"""Lane-level numpy emulation of the two-stage K_ff contraction (csrc/cov_mma.cu, TWO path: the default no-gradient K_ff kernel;
run as a CPU test by tests/test_host_logic.py::test_two_stage_contraction_emulation).

Checks, with the register layouts of mma.sync.m8n8k4.f64 on sm_100a
    A fragment: lane holds A[row = lane // 4][k = lane % 4]
    B fragment: lane holds B[k = lane % 4][n = lane // 4]
    C fragment: lane holds C[row = lane // 4][col = 2 * (lane % 4) + {0, 1}]
that (1) the accumulator registers of stage 1 can be fed, unchanged, as A fragments of stage 2 when the 8 rows b of the
column tile are taken in the k order b = 2 * (lane % 4) + j for k-step j; (2) the B fragments of stage 2 can be read from the
existing slab layout P[(kstep * 32) + row * 4 + kk] (the layout the production kernel streams with TMA); (3) the per-group
epilogue out_ce = sum_a A~_c(a) . Z_e(a) reproduces the 3x3 block of rbf_kff_many (no-gradient variant) for a pair of
8-row tiles, including a column tile that straddles two groups (stage 2 repeated per segment with masked weights).

    python profiles/experiments/two_stage_emulation.py
"""
import numpy as np

D, KS = 30, 8                       # descriptor length, k-steps of 4 (padded to 32)
LANES = np.arange(32)


def slab(mat):
    """[8, D] rows -> production slab layout [KS * 32]: element (row, col) at (col // 4) * 32 + row * 4 + col % 4."""
    out = np.zeros(KS * 32)
    for r in range(8):
        for c in range(D):
            out[(c // 4) * 32 + r * 4 + c % 4] = mat[r, c]
    return out


def mma(c, a_frag, b_frag):
    """One m8n8k4: c [32 lanes, 2] += A (8x4) . B (4x8) with per-lane fragments."""
    A = np.zeros((8, 4))
    B = np.zeros((4, 8))
    A[LANES // 4, LANES % 4] = a_frag
    B[LANES % 4, LANES // 4] = b_frag
    C = A @ B
    c[:, 0] += C[LANES // 4, 2 * (LANES % 4)]
    c[:, 1] += C[LANES // 4, 2 * (LANES % 4) + 1]


def weights(s, sigma, ell, zeta=2.0):
    """w1 = g beta, w2 = g gamma of DESIGN.md §2 (RBF, no gradient)."""
    c = 1.0 / (2 * ell * ell)
    E = sigma ** 2 * np.exp(-(1.0 - s ** zeta) * c)
    g = E * c
    beta = zeta * s ** (zeta - 1)
    gamma = zeta * (zeta - 1) * s ** (zeta - 2) + zeta ** 2 * s ** (2 * zeta - 2) * c
    return g * beta, g * gamma


def main():
    rng = np.random.default_rng(0)
    sigma, ell = 1.3, 0.7
    # row tile: 8 rows of one force group; column tile: rows 0..4 belong to group J1, rows 5..7 to group J2
    xa = rng.normal(size=(8, D)); xa /= np.linalg.norm(xa, axis=1)[:, None]
    xb = rng.normal(size=(8, D)); xb /= np.linalg.norm(xb, axis=1)[:, None]
    Aa = rng.normal(size=(3, 8, D)); Aa -= (Aa * xa).sum(-1, keepdims=True) * xa      # A~_c(a) is orthogonal to x^(a)
    Bb = rng.normal(size=(3, 8, D)); Bb -= (Bb * xb).sum(-1, keepdims=True) * xb
    seg = np.array([0, 0, 0, 0, 0, 1, 1, 1])

    # ---- reference: the 16 dot products per pair and the direct sums ---------------------------------------------
    s = xa @ xb.T
    p = np.einsum('cad,bd->cab', Aa, xb)
    q = np.einsum('ad,ebd->eab', xa, Bb)
    G = np.einsum('cad,ebd->ceab', Aa, Bb)
    w1, w2 = weights(s, sigma, ell)
    want = np.zeros((2, 3, 3))
    for J in range(2):
        m = (seg == J)[None, :]
        want[J] = np.einsum('ab,ceab->ce', w1 * m, G) + np.einsum('ab,cab,eab->ce', w2 * m, p, q)

    # ---- stage 1: x^(a) against [x^; B~_e](b): 4 accumulators, 8 k-steps each = 32 DMMAs ----------------------------
    sA = slab(xa)
    sB = [slab(xb)] + [slab(Bb[e]) for e in range(3)]
    acc = np.zeros((4, 32, 2))
    for k in range(KS):
        a_frag = sA[k * 32 + LANES]                               # lane = row * 4 + kk: one 256-byte read, as today
        for comp in range(4):
            mma(acc[comp], a_frag, sB[comp][k * 32 + LANES])     # B fragment: lane = n * 4 + kk with n = column-tile row
    s_l, q_l = acc[0], acc[1:]                                    # lane holds pairs (a = lane // 4, b = 2 * (lane % 4) + j)
    a_of, b_of = LANES // 4, 2 * (LANES % 4)
    assert np.allclose(s_l[:, 0], s[a_of, b_of]) and np.allclose(s_l[:, 1], s[a_of, b_of + 1])
    W1, W2 = weights(s_l, sigma, ell)                             # per-lane scalar epilogue, two pairs per lane

    # ---- stage 2, once per column segment: Z_e[a, :] += W1 . B~_e + (W2 q_e) . x^  (weights as A fragments) --------
    got = np.zeros((2, 3, 3))
    n_dmma = 32
    for J in range(2):
        on = np.stack([(seg[b_of] == J), (seg[b_of + 1] == J)], axis=1)
        Z = np.zeros((3, 4, 32, 2))                               # [e][n-tile][lane][2]: Z_e[a = lane//4][col = 8t + 2(lane%4) + i]
        for j in range(2):                                        # k-step j covers the column rows b = 2 * (lane % 4) + j
            a_w1 = np.where(on[:, j], W1[:, j], 0.0)              # the accumulator register IS the A fragment: no shuffle
            for e in range(3):
                a_v = np.where(on[:, j], W2[:, j] * q_l[e][:, j], 0.0)
                for t in range(4):
                    # B fragment of n-tile t: element (k = lane % 4 -> row b = 2k + j, n = lane // 4 -> col 8t + n) of the slab
                    col = 8 * t + LANES // 4
                    addr = (col // 4) * 32 + (2 * (LANES % 4) + j) * 4 + col % 4
                    mma(Z[e][t], a_w1, sB[1 + e][addr])
                    mma(Z[e][t], a_v, sB[0][addr])
                    n_dmma += 2
        # ---- per-group epilogue: out_ce = sum_a sum_col A~_c[a, col] Z_e[a, col], reduced over the lanes of the warp ----
        for c in range(3):
            sAc = slab(Aa[c])
            for e in range(3):
                tot = 0.0
                for t in range(4):
                    for i in range(2):
                        col = 8 * t + 2 * (LANES % 4) + i
                        a_val = sAc[(col // 4) * 32 + (LANES // 4) * 4 + col % 4]
                        tot += (a_val * Z[e][t][:, i]).sum()
                got[J, c, e] = tot
    err = np.abs(got - want).max() / np.abs(want).max()
    print("two-stage vs direct: max rel err %.2e ; DMMAs for this straddling tile: %d (one segment: 80, 4x4 block kernel: 128)"
          % (err, n_dmma))
    # bank picture of the stage-2 B fragment read (8-byte words, 16 words per 128-byte wavefront line)
    col = 8 * 0 + LANES // 4
    addr = (col // 4) * 32 + (2 * (LANES % 4)) * 4 + col % 4
    banks = len(set(addr % 16))
    print("stage-2 B fragment: %d distinct 8-byte banks for 32 lanes -> %d wavefronts (a conflict-free 256-byte read takes 2)"
          % (banks, 32 // banks))
    assert err < 1e-12


if __name__ == "__main__":
    main()
