"""TEST INFRASTRUCTURE -- numpy restatement of the integer arithmetic behind the tcgen05 trailing
update of the blocked reduced solve (pysfm_b200/csrc/ba_solve_tc.cuh).  Only tests/ import it.

What it restates.  The reference solves the reduced camera system with numpy.linalg.solve
(bundle_adjuster.py:303); the product factors large systems panel by panel and applies each panel
to the trailing matrix,  A22 -= L21 L21^T,  as INT8 tensor-core products (tcgen05 has no FP64 kind).
This file is that one step in plain numpy, operation for operation:

  slice_rows       row scale 2^e_i with |L_ik| 2^-e_i < 1, then S signed 7-bit digits per entry,
                   x = sum_p d_p 2^(-6-7p) + O(2^-7S), every step exact in FP64  (slice_panel_kernel)
  level_sums       for level l = p + q:  sum_k d_p[i,k] d_q[j,k]  in exact integers  (UTCIMMA into TMEM)
  combine          Horner in 2^-7 from the smallest level up, then the two row scales   (epilogue)
  trailing_update  the three together on a dense matrix, plus the right-hand side update

The device results are held to these BIT FOR BIT (digits, scales, INT32 level sums, updated
matrix) in tests/test_gpu_tc_solve.py; tests/test_ozaki_model.py checks the model itself against
FP64 (error bound 2^-7S) and inside a blocked Cholesky of real reduced camera systems built by
oracle/ba_oracle.py (which IS pinned to the reference).
"""
import numpy as np


def slice_rows(X, S):
    """X (m, K) float64 -> (e (m,) int, digits (S, m, K) int8) with
    X[i, k] = 2^e_i * (sum_p digits[p, i, k] * 2^(-6-7p) + r),  |r| <= 2^(-7S)."""
    X = np.asarray(X, dtype=np.float64)
    mu = np.abs(X).max(axis=1)
    e = np.where(mu > 0, np.frexp(mu)[1], 0).astype(np.int64)   # mu = f 2^e, f in [0.5, 1)
    t = np.ldexp(X, (6 - e)[:, None].astype(np.int32))          # x * 64, |t| < 64
    digits = np.empty((S,) + X.shape, dtype=np.int8)
    for p in range(S):
        d = np.rint(t)                                          # round half to even, like rint() on the device
        digits[p] = d.astype(np.int8)
        t = (t - d) * 128.0                                     # exact
    return e, digits


def level_sums(da, db, S):
    """da (S, m, K), db (S, n, K) int8 -> (S, m, n) int64: sum over p+q = l, k of da[p,i,k] db[q,j,k]."""
    da = da.astype(np.int64)
    db = db.astype(np.int64)
    out = np.zeros((S, da.shape[1], db.shape[1]), dtype=np.int64)
    for l in range(S):
        for p in range(l + 1):
            out[l] += da[p] @ db[l - p].T
    assert np.abs(out).max(initial=0) < 2 ** 31                  # what the INT32 accumulators rely on
    return out


def combine(sums, scale_rows, scale_cols):
    """(S, m, n) level sums -> FP64 product: Horner in 2^-7, then the power-of-two scales 2^(e-6)."""
    S = sums.shape[0]
    v = np.zeros(sums.shape[1:], dtype=np.float64)
    for l in range(S - 1, -1, -1):
        v = v * 0.0078125 + sums[l].astype(np.float64)
    return v * (scale_rows[:, None] * scale_cols[None, :])


def syrk(X, S):
    """X X^T through S slices (what the tensor cores + epilogue compute for one panel)."""
    e, d = slice_rows(X, S)
    sc = np.ldexp(1.0, (e - 6).astype(np.int32))
    return combine(level_sums(d, d, S), sc, sc)


def trailing_update(A, rhs, K, S, saved_rhs=None):
    """The device step on a dense (ld, ld) matrix: panel = A[K:, :K]; lower triangle of A[K:, K:] and
    rhs[K:] updated.  Returns dict(A, rhs, e, digits, scale, level_sums) (level sums for the trailing
    block only, (S, ld-K, ld-K))."""
    A = np.array(A, dtype=np.float64)
    rhs = np.array(rhs, dtype=np.float64)
    L = A[K:, :K]
    e, d = slice_rows(L, S)
    sc = np.ldexp(1.0, (e - 6).astype(np.int32))
    ls = level_sums(d, d, S)
    upd = combine(ls, sc, sc)
    low = np.tril(np.ones(upd.shape, dtype=bool))
    A[K:, K:] = np.where(low, A[K:, K:] - upd, A[K:, K:])
    b = rhs[K:].copy()
    if saved_rhs is not None:
        b[:64] = saved_rhs
    rhs[K:] = b - L @ rhs[:K]
    return dict(A=A, rhs=rhs, e=e, digits=d, scale=sc, level_sums=ls)


def blocked_cholesky_solve(A, b, nb, S):
    """Right-looking blocked Cholesky solve whose trailing updates go through `syrk` (S slices;
    S = None: plain FP64) -- the numerical skeleton of ba_solve's blocked path."""
    A = np.array(A, dtype=np.float64)
    n = A.shape[0]
    L = np.zeros_like(A)
    for c0 in range(0, n, nb):
        c1 = min(n, c0 + nb)
        L11 = np.linalg.cholesky(A[c0:c1, c0:c1])
        L[c0:c1, c0:c1] = L11
        if c1 < n:
            L21 = np.linalg.solve(L11, A[c1:, c0:c1].T).T
            L[c1:, c0:c1] = L21
            A[c1:, c1:] -= (L21 @ L21.T) if S is None else syrk(L21, S)
    y = np.linalg.solve(L, b)
    return np.linalg.solve(L.T, y)
