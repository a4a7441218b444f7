"""Host emulation of the tile algorithm behind DCGP_LEAF_FINV (csrc/chol.cu: block_potrf<T, true> with inv_rows_init /
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


def diag_step_lanes(block, pw, piv2):
    """diag_step of csrc/chol.cu lane by lane (lane r holds row r of the 16 x 16 block; entries above the diagonal are
    junk that is never stored): `piv2` broadcasts the next pivot from `row[c+1] - lrc^2` of lane c+1 instead of the
    generally updated row[c+1] (DCGP_LEAF_PIV2).  Returns the factor and the inverse pivots."""
    lanes = 32
    row = np.zeros((lanes, SB))
    for r in range(lanes):
        for c in range(SB):
            row[r, c] = block[r, c] if (r < pw and c <= r) else (1.0 if r == c else 0.0)
    myinv = np.zeros(lanes)
    dnext = row[0, 0]
    for c in range(SB):
        d = dnext if piv2 else row[c, c]                      # shuffle from lane c
        y = 1.0 / np.sqrt(d)
        scaled = row[:, c] * y
        for r in range(lanes):
            row[r, c] = d * y if r == c else scaled[r]
        myinv[c] = y
        lrc = row[:, c].copy()
        if piv2 and c + 1 < SB:
            own = -lrc * lrc + row[:, c + 1]
            dnext = own[c + 1]                                # shuffle from lane c + 1
        for k in range(c + 1, SB):
            lkc = row[k, c]                                   # shuffle of row[c] from lane k
            row[:, k] = -lrc * lkc + row[:, k]
    L = np.zeros((pw, pw))
    for r in range(pw):
        L[r, :r + 1] = row[r, :r + 1]
    return L, myinv[:pw]


@pytest.mark.parametrize("pw", [16, 9, 1])
def test_short_pivot_chain_is_the_same_arithmetic(pw):
    rng = np.random.default_rng(pw)
    M = rng.standard_normal((pw, pw))
    A = M @ M.T + pw * np.eye(pw)
    blk = np.zeros((32, SB))
    blk[:pw, :pw] = np.tril(A)
    L0, i0 = diag_step_lanes(blk, pw, False)
    L1, i1 = diag_step_lanes(blk, pw, True)
    assert np.array_equal(L0, L1) and np.array_equal(i0, i1)          # bitwise: same operands, same order
    assert np.abs(L0 - np.linalg.cholesky(A)).max() < 1e-12 * np.abs(A).max()
