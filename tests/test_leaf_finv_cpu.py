"""Host emulation of the tile algorithm of the leaves (DCGP_LEAF_FINV, default since round 2) (csrc/chol.cu: block_potrf<T, true> with inv_rows_init /
inv_rows_panel / inv_rows_update): right-looking Cholesky over 16-wide sub-panels in a pitch-129 tile whose unused upper
part carries the unit rows e_c, so that L^{-1} is complete when the last pivot is ( X(i, c) <-> S[c][i + 1] ).  Same
index expressions and predicates as the kernel, executed sequentially; the tile starts filled with NaN, so any read of
an entry the kernel may not read shows up in the result."""
import numpy as np
import pytest

NB, SB = 128, 16
SP = NB + 1


def fused_leaf(A):
    nb = A.shape[0]
    S = np.full((NB, SP), np.nan)
    for i in range(nb):
        S[i, :i + 1] = A[i, :i + 1]
    rinv = np.zeros(NB)

    def diag_step(p0, pw):
        blk = np.zeros((pw, pw))
        for r in range(pw):
            blk[r, :r + 1] = S[p0 + r, p0:p0 + r + 1]
        L = np.linalg.cholesky(blk + np.tril(blk, -1).T)
        for r in range(pw):
            S[p0 + r, p0:p0 + r + 1] = L[r, :r + 1]
            rinv[p0 + r] = 1.0 / L[r, r]

    for e in range(NB * NB):                                   # inv_rows_init
        c, k = e >> 7, e & (NB - 1)
        if c < nb and k < nb and k >= c:
            S[c, k + 1] = 1.0 if k == c else 0.0
    diag_step(0, min(SB, nb))
    for p0 in range(0, nb, SB):
        pw = min(SB, nb - p0)
        for i in range(p0 + pw, nb):                           # step (2), rows below
            x = [S[i, p0 + c] if c < pw else 0.0 for c in range(SB)]
            for c in range(pw):
                x[c] = (x[c] - sum(x[k] * S[p0 + c, p0 + k] for k in range(c))) * rinv[p0 + c]
            S[i, p0:p0 + pw] = x[:pw]
        for c in range(p0 + pw):                               # inv_rows_panel
            x = [S[c, p0 + kk + 1] if (kk < pw and p0 + kk >= c) else 0.0 for kk in range(SB)]
            for kk in range(pw):
                x[kk] = (x[kk] - sum(x[k] * S[p0 + kk, p0 + k] for k in range(kk))) * rinv[p0 + kk]
            for kk in range(pw):
                if p0 + kk >= c:
                    S[c, p0 + kk + 1] = x[kk]
        t0 = p0 + pw
        if t0 >= nb:
            break
        P = S[t0:nb, p0:t0].copy()
        upd = P @ P.T                                          # trailing_update (lower part only)
        for row in range(t0, nb):
            S[row, t0:row + 1] -= upd[row - t0, :row - t0 + 1]
        for c in range(t0):                                    # inv_rows_update
            a = np.array([S[c, p0 + k + 1] if p0 + k >= c else 0.0 for k in range(pw)])
            S[c, t0 + 1:nb + 1] -= P @ a
        diag_step(t0, min(SB, nb - t0))
    L = np.zeros((nb, nb))
    X = np.zeros((nb, nb))
    for i in range(nb):
        L[i, :i + 1] = S[i, :i + 1]
        X[i, :i + 1] = [S[c, i + 1] for c in range(i + 1)]
    return L, X


@pytest.mark.parametrize("nb", [128, 113, 100, 33, 16, 7, 1])
def test_fused_leaf_gives_factor_and_inverse(nb):
    rng = np.random.default_rng(nb)
    M = rng.standard_normal((nb, nb))
    A = M @ M.T + nb * np.eye(nb)
    L, X = fused_leaf(A)
    ref = np.linalg.cholesky(A)
    assert np.isfinite(L).all() and np.isfinite(X).all()
    assert np.abs(L - ref).max() < 1e-12 * np.abs(ref).max()
    assert np.abs(X @ ref - np.eye(nb)).max() < 1e-12


def diag_pair_lanes(block, pw, dtype=np.float64):
    """FP32 diag_step of csrc/chol.cu lane by lane (lane r holds row r of the 16 x 16 block, identity padding outside
    the valid pw x pw part): two columns per round, 1/sqrt(a) and 1/sqrt(a e - b^2) from the same three broadcasts,
    second pivot's reciprocal root = sqrt(a) / sqrt(det).  Returns the factor, the inverse pivots and `bad`."""
    lanes = 32
    row = np.zeros((lanes, SB), dtype=dtype)
    for r in range(lanes):
        for c in range(SB):
            row[r, c] = block[r, c] if (r < pw and c <= r) else (1.0 if r == c else 0.0)
    myinv = np.zeros(lanes, dtype=dtype)
    bad = 0
    one = dtype(1.0)
    for c in range(0, SB, 2):
        a, b, e = row[c, c], row[c + 1, c], row[c + 1, c + 1]       # three shuffles
        det = a * e - b * b
        if bad == 0:
            if not a > 0:
                bad = c + 1
            elif not det > 0:
                bad = c + 2
        y1, yd = one / np.sqrt(a), one / np.sqrt(det)
        sa = a * y1
        y2 = sa * yd
        l21 = b * y1
        l1 = row[:, c] * y1
        l1[c] = sa
        t = row[:, c + 1] - l1 * l21
        l2 = t * y2
        l2[c + 1] = (det * yd) * y1
        row[:, c], row[:, c + 1] = l1, l2
        myinv[c], myinv[c + 1] = y1, y2
        for k in range(c + 2, SB):
            row[:, k] = row[:, k] - l1 * l1[k] - l2 * l2[k]           # shuffles of l1, l2 from lane k
    L = np.zeros((pw, pw), dtype=dtype)
    for r in range(pw):
        L[r, :r + 1] = row[r, :r + 1]
    return L, myinv[:pw], bad


@pytest.mark.parametrize("pw", [16, 15, 9, 2, 1])
def test_two_columns_per_round_pivots_give_the_cholesky_factor(pw):
    rng = np.random.default_rng(pw)
    M = rng.standard_normal((pw, pw))
    A = M @ M.T + pw * np.eye(pw)
    blk = np.zeros((32, SB))
    blk[:pw, :pw] = np.tril(A)
    L, inv, bad = diag_pair_lanes(blk, pw)
    ref = np.linalg.cholesky(A)
    assert bad == 0
    assert np.abs(L - ref).max() < 1e-13 * np.abs(A).max()
    assert np.abs(inv * np.diag(ref) - 1).max() < 1e-13
    L32, _, _ = diag_pair_lanes(blk.astype(np.float32), pw, np.float32)
    assert np.abs(L32 - ref).max() < 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("col", [0, 1, 6, 7, 15])
def test_two_columns_per_round_pivots_report_the_first_bad_minor(col):
    """LAPACK-style info: the first non-positive leading minor, whether it is the first or the second column of a pair."""
    rng = np.random.default_rng(col)
    M = rng.standard_normal((16, 16))
    A = M @ M.T + 16 * np.eye(16)
    L = np.linalg.cholesky(A)
    L[col, col] = 0.0                      # A = L L^T is then singular exactly at leading minor col + 1
    A = L @ L.T
    A[col, col] -= 1e-3
    blk = np.zeros((32, SB))
    blk[:16, :16] = np.tril(A)
    with np.errstate(invalid="ignore", divide="ignore"):
        _, _, bad = diag_pair_lanes(blk, 16)
    assert bad == col + 1
