"""CPU checks of oracle/ozaki_model.py, the numpy restatement of the tcgen05 trailing update
(pysfm_b200/csrc/ba_solve_tc.cuh): the digit decomposition is exact, the level sums fit INT32,
the combined product is within 2^-7S of FP64, and a blocked Cholesky built on it solves real
reduced camera systems (from the reference-pinned oracle) to the accuracy DESIGN.md quotes."""
import numpy as np
import pytest

from oracle import ba_oracle, ozaki_model
from pysfm_b200 import synthetic


def test_digits_reconstruct_the_entries_exactly_up_to_the_last_slice():
    rng = np.random.RandomState(0)
    X = rng.randn(37, 128) * np.ldexp(1.0, rng.randint(-30, 30, size=(37, 1)))
    X[5] = 0.0
    X[9, 3] = 12345.678
    for S in (4, 5, 6, 7):
        e, d = ozaki_model.slice_rows(X, S)
        assert d.dtype == np.int8 and np.abs(d.astype(int)).max() <= 64
        rec = sum(d[p].astype(np.float64) * np.ldexp(1.0, -6 - 7 * p) for p in range(S))
        err = np.abs(np.ldexp(X, -e[:, None].astype(np.int32)) - rec)
        assert err.max() <= np.ldexp(1.0, -7 * S) * (1 + 1e-12)
        assert (np.abs(np.ldexp(X, -e[:, None].astype(np.int32))) < 1.0).all()
    assert (d[:, 5] == 0).all()


def test_level_sums_fit_int32_at_the_largest_panel():
    # worst case: every digit +-64, K = 1024 (window 16), S = 7 pairs on the last level
    d = np.full((7, 2, 1024), 64, dtype=np.int8)
    s = ozaki_model.level_sums(d, d, 7)
    assert s.max() == 7 * 1024 * 64 * 64 and s.max() < 2 ** 31


@pytest.mark.parametrize("S", [4, 5, 6, 7])
def test_sliced_product_is_within_two_to_the_minus_7S_of_fp64(S):
    rng = np.random.RandomState(S)
    X = rng.randn(96, 256) * np.ldexp(1.0, rng.randint(-8, 8, size=(96, 1)))
    e, _ = ozaki_model.slice_rows(X, S)
    got = ozaki_model.syrk(X, S)
    want = X @ X.T
    bound = np.ldexp(1.0, (e[:, None] + e[None, :]).astype(np.int32)) * 256 * (S + 2) * np.ldexp(1.0, -7 * S)
    assert (np.abs(got - want) <= bound + 1e-13 * np.abs(want)).all()


def _reduced_system(n_cam, n_pt, k, seed, damping):
    a = synthetic.make_arrays(n_cam, n_pt, k, seed)
    P = ba_oracle.Problem(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"],
                          ('gaussian', np.eye(2)), np.arange(1, n_cam), np.arange(n_pt))
    blocks = ba_oracle.prepare(P)
    ba_oracle.apply_damping(blocks, damping)
    Sm, b, _ = ba_oracle.schur(P, blocks)
    n = 6 * Sm.shape[0]
    return Sm.transpose(0, 2, 1, 3).reshape(n, n), b.reshape(-1)


@pytest.mark.parametrize("damping,tol6", [(10.0, 1e-13), (1e-2, 1e-10), (1e-4, 1e-9)])
def test_blocked_cholesky_on_reduced_camera_systems(damping, tol6):
    """The update dC of compute_update (bundle_adjuster.py:176-208) through the blocked solve:
    6 slices stay far inside north_star's 1e-6; 4 slices are the documented floor."""
    A, b = _reduced_system(60, 1500, 8, 3, damping)
    ref = np.linalg.solve(A, b)
    scale = np.abs(ref).max()
    x6 = ozaki_model.blocked_cholesky_solve(A, b, 128, 6)
    x4 = ozaki_model.blocked_cholesky_solve(A, b, 128, 4)
    x0 = ozaki_model.blocked_cholesky_solve(A, b, 128, None)
    assert np.abs(x0 - ref).max() / scale < 1e-12
    assert np.abs(x6 - ref).max() / scale < tol6
    assert np.abs(x4 - ref).max() / scale < 1e-6


def test_trailing_update_matches_plain_fp64_and_leaves_the_rest_alone():
    rng = np.random.RandomState(11)
    ld, K = 256, 128
    A = rng.randn(ld, ld)
    rhs = rng.randn(ld)
    saved = rng.randn(64)
    out = ozaki_model.trailing_update(A, rhs, K, 6, saved_rhs=saved)
    L = A[K:, :K]
    want = A[K:, K:] - L @ L.T
    low = np.tril(np.ones((ld - K, ld - K), dtype=bool))
    assert np.abs(out["A"][K:, K:][low] - want[low]).max() < 1e-9
    assert (out["A"][K:, K:][~low] == A[K:, K:][~low]).all()
    assert (out["A"][:K] == A[:K]).all() and (out["A"][:, :K] == A[:, :K]).all()
    b = rhs[K:].copy()
    b[:64] = saved
    assert np.abs(out["rhs"][K:] - (b - L @ rhs[:K])).max() < 1e-12
    assert (out["rhs"][:K] == rhs[:K]).all()
