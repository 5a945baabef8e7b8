"""The blocked reduced solve whose trailing updates run on tcgen05 (pysfm_b200/csrc/ba_solve_tc.cuh),
through the C ABI on a B200:

  * one trailing update against its numpy restatement (oracle/ozaki_model.py) BIT FOR BIT: INT8 digit
    planes, power-of-two row scales, the INT32 level sums read back from tensor memory, the updated
    matrix; the right-hand side update to rounding;
  * solve_motion_normal_eqns (bundle_adjuster.py:281-312) on random SPD systems with frozen
    parameters, every window width / slice count, against numpy.linalg.solve;
  * BASELINE config 5 (500 cameras): compute_update through the blocked path against the oracle's
    committed steps at lambda = 1e-4 .. 1e2;
  * an indefinite system is still reported as NormalEquationsIllconditioned.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ld,window,slices,bk", [(704, 2, 6, 64), (640, 2, 7, 128), (832, 4, 4, 64), (576, 6, 5, 64)])
def test_trailing_update_bit_for_bit(ld, window, slices, bk, cuda_device):
    from oracle import ozaki_model
    from pysfm_b200 import solver_tc
    import torch
    dev = torch.device(cuda_device).index or 0
    rs = np.random.RandomState(ld + slices)
    K = 64 * window
    A = rs.randn(ld, ld)
    # panel rows of very different magnitude, a zero row, a row dominated by one entry
    A[K:, :K] *= np.ldexp(1.0, rs.randint(-12, 12, size=(ld - K, 1)))
    A[K + 5, :K] = 0.0
    A[K + 9, 3] = 12345.678
    rhs = rs.randn(ld)
    saved = rs.randn(64)
    got = solver_tc.trailing_update(A, rhs, window, slices=slices, bk=bk, saved_rhs=saved, device=dev)
    want = ozaki_model.trailing_update(A, rhs, K, slices, saved_rhs=saved)
    m = ld - K
    assert (got["digits"][:, K:ld, :] == want["digits"]).all()
    assert (got["scale"][K:ld] == want["scale"]).all()
    low = np.tril(np.ones((m, m), dtype=bool))
    dev_sums = got["level_sums"][:, K:ld, K:ld]
    assert (dev_sums[:, low] == want["level_sums"][:, low]).all()
    assert (got["A"] == want["A"]).all()                      # the updated lower triangle AND everything left alone
    assert relerr(got["rhs"], want["rhs"]) < 1e-13
    # and the point of it all: the update is the FP64 product to 2^-7S of the row scales
    L = A[K:, :K]
    err = np.abs((A[K:, K:] - got["A"][K:, K:]) - L @ L.T)[low].max()
    bound = (want["scale"].max() * 64) ** 2 * K * (slices + 2) * 2.0 ** (-7 * slices)
    assert err <= bound + 1e-12 * np.abs(L @ L.T).max()


@pytest.mark.parametrize("nc,window,slices", [(200, 8, 6), (200, 2, 7), (200, 4, 5), (500, 8, 6), (500, 16, 6), (333, 6, 6)])
def test_blocked_solve_random_spd_systems(nc, window, slices, cuda_device):
    from pysfm_b200 import synthetic, _lib
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    b = synthetic.make_scene(nc + 1, 40, 4, 31)
    ba = BundleAdjuster(b, device=cuda_device, verbose=False)
    p = ba._problem
    p.set_option(_lib.BA_OPT_TC_MIN_TILES, 1)
    p.set_option(_lib.BA_OPT_TC_WINDOW, window)
    p.set_option(_lib.BA_OPT_TC_SLICES, slices)
    rs = np.random.RandomState(7 * nc + window)
    n = 6 * nc
    G = rs.randn(n, n + 5)
    A = G @ G.T / n + np.diag(rs.rand(n) + 0.5)
    rhs = rs.randn(n)
    S = A.reshape(nc, 6, nc, 6).transpose(0, 2, 1, 3).copy()
    tol = {7: 1e-9, 6: 1e-9, 5: 1e-7}[slices]
    launches0 = p.launch_count()
    x = ba.solve_motion_normal_eqns(S, rhs.reshape(nc, 6), np.ones(n, bool))
    assert p.launch_count() - launches0 > 6            # the blocked path ran (expand + dataflow kernel alone are 2 launches)
    assert relerr(x.flatten(), np.linalg.solve(A, rhs)) < tol
    mask = rs.rand(n) > 0.3
    mask[0] = True
    x = ba.solve_motion_normal_eqns(S, rhs.reshape(nc, 6), mask)
    ref = np.zeros(n)
    ref[mask] = np.linalg.solve(A[mask][:, mask], rhs[mask])
    assert relerr(x.flatten(), ref) < tol
    # the FP64 dataflow solve of the same system, for the record of what the slices cost
    p.set_option(_lib.BA_OPT_TC_MIN_TILES, 0)
    x0 = ba.solve_motion_normal_eqns(S, rhs.reshape(nc, 6), mask)
    assert relerr(x.flatten(), x0.flatten()) < tol


def test_blocked_solve_reports_an_indefinite_system(cuda_device):
    from pysfm_b200 import synthetic, _lib
    from pysfm_b200.bundle_adjuster import BundleAdjuster, NormalEquationsIllconditioned
    nc = 200
    b = synthetic.make_scene(nc + 1, 40, 4, 31)
    ba = BundleAdjuster(b, device=cuda_device, verbose=False)
    ba._problem.set_option(_lib.BA_OPT_TC_MIN_TILES, 1)
    rs = np.random.RandomState(5)
    n = 6 * nc
    G = rs.randn(n, n)
    A = G @ G.T / n + np.eye(n)
    A[700, 700] = -3.0                                   # a negative pivot in the second window
    S = A.reshape(nc, 6, nc, 6).transpose(0, 2, 1, 3).copy()
    with pytest.raises(NormalEquationsIllconditioned):
        ba.solve_motion_normal_eqns(S, rs.randn(nc, 6), np.ones(n, bool))


def test_config5_steps_through_the_blocked_solve(cuda_device):
    """500 cameras / 200,000 points / 2 M observations, lambda = 1e-4 .. 1e2: the whole LM step with the
    reduced system (2,994 unknowns, 47 tiles) factored by the blocked tcgen05 path, against the oracle."""
    path = os.path.join(GOLDEN, "config5_sweep.npz")
    if not os.path.isfile(path):
        pytest.skip("tests/golden/config5_sweep.npz not generated (oracle/make_golden_large.py c5)")
    from pysfm_b200 import synthetic, _lib
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    g = load_golden("config5_sweep")
    b = synthetic.make_config("C5")
    ba = BundleAdjuster(b, device=cuda_device, verbose=False)
    ba._problem.set_option(_lib.BA_OPT_TC_MIN_TILES, 1)
    stride = int(g["sample_stride"])
    for i, lam in enumerate(g["lambdas"]):
        n0 = ba._problem.launch_count()
        motion, structure = ba.compute_update(float(lam))
        assert ba._problem.launch_count() - n0 > 12
        _, cand, status = ba._problem.read_scalars()
        assert status == 0
        assert relerr(motion, g["sweep_motion"][i]) < 1e-7, lam
        assert relerr(structure[::stride], g["sweep_structure_sample"][i]) < 1e-7, lam
        assert abs(cand - float(g["sweep_cand_cost"][i])) < 1e-8 * float(g["sweep_cand_cost"][i]), lam


def test_blocked_solve_event_breakdown(cuda_device):
    """ba_tc_solve_profile: CUDA-event times of the stages of the blocked solve (stage timers)."""
    from pysfm_b200 import synthetic, _lib
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    nc = 300
    b = synthetic.make_scene(nc + 1, 40, 4, 31)
    ba = BundleAdjuster(b, device=cuda_device, verbose=False)
    p = ba._problem
    p.set_option(_lib.BA_OPT_TC_MIN_TILES, 1)
    p.set_option(_lib.BA_OPT_SOLVER_PROFILE, 1)
    rs = np.random.RandomState(3)
    n = 6 * nc
    G = rs.randn(n, n + 5)
    A = G @ G.T / n + np.eye(n)
    S = A.reshape(nc, 6, nc, 6).transpose(0, 2, 1, 3).copy()
    rhs = rs.randn(nc, 6)
    p.tc_solve_profile(reset=True)
    x = ba.solve_motion_normal_eqns(S, rhs, np.ones(n, bool))
    x = ba.solve_motion_normal_eqns(S, rhs, np.ones(n, bool))
    prof = p.tc_solve_profile(reset=True)
    assert prof["solves"] == 2
    for k in ("expand", "panels", "slices", "trailing_updates", "backward"):
        assert 0.0 < prof[k] < 50.0, (k, prof)
    assert p.tc_solve_profile()["solves"] == 0
    assert relerr(x.flatten(), np.linalg.solve(A, rhs.reshape(-1))) < 1e-9
