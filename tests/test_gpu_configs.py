"""BASELINE.json configurations at FULL size on one B200, CUDA path through the C ABI against the
CPU oracle -- run live where the oracle finishes in seconds (config 2), against its committed
outputs where it takes minutes (config 5, oracle/make_golden_large.py):

  config 2   direct parity of compute_update, cost and the residual vector;
  config 5   one LM step at lambda = 1e-4 .. 1e2 and the free-running optimize() curve;
  configs 1/3  reprojection RMSE of the optimised bundle against the unmodified reference's;
plus the robustness switches this round added (strict flag publication, spin-wait deadline,
non-finite costs), which must not change a single bit / must report instead of hanging.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, golden_bundle, relerr

pytestmark = pytest.mark.gpu


def _oracle_problem(a):
    from oracle import ba_oracle
    nc, nt = len(a["Rs"]), len(a["pts"])
    return ba_oracle.Problem(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"],
                             ('gaussian', np.eye(2)), np.arange(1, nc), np.arange(nt))


def _rmse(bundle):
    """sqrt(mean ||pred - z||^2) in pixels over every measurement of an array-backed bundle (the
    oracle's formula applied to the bundle's parameters: a checker, not the product)."""
    from oracle import ba_oracle
    o_cam, o_trk, o_uv = bundle._obs_arrays
    P = ba_oracle.Problem(bundle.K, bundle.Rs(), bundle.ts(), bundle.reconstruction, o_cam, o_trk, o_uv,
                          ('gaussian', np.eye(2)), np.arange(1, len(bundle.cameras)), np.arange(bundle.num_tracks()))
    return ba_oracle.reprojection_rmse(P)


def test_config2_direct_parity_with_the_oracle(cuda_device):
    """200 cameras / 50,000 points / 500,000 observations: update, costs and residuals <= 1e-6
    relative (north_star); measured here at ~1e-13."""
    from oracle import ba_oracle
    from pysfm_b200 import synthetic
    from pysfm_b200.bundle import Bundle
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    cfg = synthetic.CONFIGS["C2"]
    a = synthetic.make_arrays(**cfg)
    b = Bundle.FromObservationArrays(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"])
    ba = BundleAdjuster(b, device=cuda_device, verbose=False)
    motion, structure = ba.compute_update(10.0)
    cost, cand, status = ba._problem.read_scalars()
    P = _oracle_problem(a)
    m2, s2 = ba_oracle.compute_update(P, 10.0)
    assert status == 0
    assert relerr(motion, m2) < 1e-8
    assert relerr(structure, s2) < 1e-8
    c0 = ba_oracle.compute_cost(P)
    assert abs(cost - c0) < 1e-11 * c0
    c1 = ba_oracle.compute_cost(ba_oracle.apply_update(P, m2, s2))
    assert abs(cand - c1) < 1e-9 * c1
    # the residual vector itself, observation by observation (Bundle.residuals enumerates tracks in
    # order and a track's cameras in its own order -- the generator's, and the oracle's, order)
    r, _, _ = ba_oracle.linearize(P)
    assert relerr(np.asarray(b.residuals()).reshape(-1, 2), r) < 1e-9


def test_config5_damping_sweep_and_convergence_curve(cuda_device):
    """500 cameras / 200,000 points / 2 M observations: single steps at lambda = 1e-4 .. 1e2 and the whole
    optimize() trace (every trial's lambda, cost, candidate cost, accept/reject) against the oracle."""
    path = os.path.join(GOLDEN, "config5_sweep.npz")
    if not os.path.isfile(path):
        pytest.skip("tests/golden/config5_sweep.npz not generated (oracle/make_golden_large.py c5)")
    from pysfm_b200 import synthetic
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    g = load_golden("config5_sweep")
    b = synthetic.make_config("C5")
    ba = BundleAdjuster(b, device=cuda_device, verbose=False)
    stride = int(g["sample_stride"])
    c0 = ba.compute_cost(b)
    assert abs(c0 - float(g["cost0"])) < 1e-11 * float(g["cost0"])
    assert abs(_rmse(b) - float(g["rmse0"])) < 1e-9
    for i, lam in enumerate(g["lambdas"]):
        motion, structure = ba.compute_update(float(lam))
        _, cand, status = ba._problem.read_scalars()
        assert status == 0
        assert relerr(motion, g["sweep_motion"][i]) < 1e-7, lam
        assert relerr(structure[::stride], g["sweep_structure_sample"][i]) < 1e-7, lam
        assert abs(float(np.sum(structure * structure)) - float(g["sweep_structure_sq"][i])) < 1e-7 * float(g["sweep_structure_sq"][i])
        assert abs(cand - float(g["sweep_cand_cost"][i])) < 1e-8 * float(g["sweep_cand_cost"][i]), lam
    ba.optimize(max_steps=25)
    ref = g["opt_costs"]
    assert len(ba.costs) == len(ref) and ba.num_steps == int(g["opt_num_steps"]) and ba.converged == bool(g["opt_converged"])
    assert relerr(np.array(ba.costs), ref) < 1e-8
    assert [t["damping"] for t in ba.trace] == [float(x) for x in g["opt_trace_damping"]]
    assert [bool(t["accepted"]) for t in ba.trace] == [bool(x) for x in g["opt_trace_accepted"]]
    assert relerr(np.array([t["cand_cost"] for t in ba.trace]), g["opt_trace_cand_cost"]) < 1e-8
    assert abs(_rmse(ba.bundle) - float(g["opt_rmse_final"])) < 1e-7
    assert relerr(ba.bundle.reconstruction[::stride], g["opt_pts_sample"]) < 1e-6
    assert relerr(ba.bundle.Rs(), g["opt_Rs"]) < 1e-6


@pytest.mark.parametrize("name", ["config1_synthetic", "oleg_synthetic"])
def test_reprojection_rmse_against_the_reference(name, cuda_device):
    """BASELINE metric, second half: reprojection RMSE of the optimised bundle vs the UNMODIFIED
    reference's (configs 1 and 3; fixtures written by oracle/make_golden.py through refshim)."""
    from pysfm_b200.bundle import Bundle
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    g = load_golden(name)
    b = golden_bundle(g)
    ba = BundleAdjuster(b, device=cuda_device, verbose=False)
    ba.optimize(max_steps=int(g["opt_num_steps"]))
    ref = Bundle.FromObservationArrays(g["K"], g["opt_Rs"], g["opt_ts"], g["opt_pts"], g["obs_cam"].astype(np.int64),
                                       g["obs_track"].astype(np.int64), g["obs_uv"].astype(np.float64),
                                       sensor_model=b.sensor_model)
    rmse_ref, rmse_ours, rmse_init = _rmse(ref), _rmse(ba.bundle), _rmse(b)
    assert rmse_ours < rmse_init
    assert abs(rmse_ours - rmse_ref) < 1e-6 * max(rmse_ref, 1.0)
    assert abs(ba.bundle.complete_cost() - float(g["opt_complete_cost"])) < 1e-6 * float(g["opt_complete_cost"])


def test_strict_flag_publication_gives_the_same_bits(cuda_device):
    """The solver publishes column-block flags with relaxed stores behind completed bulk copies
    (DESIGN.md 4.2); BA_OPT_STRICT_FLAGS switches to release/acquire.  Same system, many solves,
    both ways: every solution must be bit-identical (a lost ordering would show up as a different
    -- stale -- operand in some tile product)."""
    from pysfm_b200 import _lib
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    from pysfm_b200 import synthetic
    b = synthetic.make_scene(200, 3000, 10, seed=31)
    ba = BundleAdjuster(b, device=cuda_device, verbose=False)
    p = ba._problem
    ba._push(b)
    p.linearize_eliminate(10.0, 1e-5, _lib.BA_WANT_SCHUR)     # one system, factored over and over
    ref = None
    for strict in (0, 1, 0, 1):
        p.set_option(_lib.BA_OPT_STRICT_FLAGS, strict)
        for _ in range(60):
            p.solve(None)
            x = p.get_array(_lib.BA_ARR_DC, (p.n_sys,))
            assert p.read_scalars()[2] == 0
            if ref is None:
                ref = x
            assert np.array_equal(x, ref), "solver result depends on timing (strict=%d)" % strict


def test_spin_wait_deadline_reports_instead_of_hanging(cuda_device):
    """A solver launched with fewer CTAs than it needs cannot happen in the product (the grid is
    capped at the SM count and tasks are handed out in dependency order), so the deadline is
    exercised the other way round: a budget of a microsecond makes the first idle poll give up;
    the launch must END and report BA_ERR_TIMEOUT."""
    from pysfm_b200 import _lib
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    from pysfm_b200 import synthetic
    b = synthetic.make_scene(120, 2000, 8, seed=32)
    ba = BundleAdjuster(b, device=cuda_device, verbose=False)
    p = ba._problem
    ba._push(b)
    p.linearize_eliminate(10.0, 1e-5, _lib.BA_WANT_SCHUR)
    p.set_option(_lib.BA_OPT_SPIN_TIMEOUT_MS, 1e-6)
    p.solve(None)
    assert p.read_scalars()[2] == _lib.BA_ERR_TIMEOUT
    p.set_option(_lib.BA_OPT_SPIN_TIMEOUT_MS, 10000.0)
    p.solve(None)
    assert p.read_scalars()[2] == 0
    m1 = p.get_array(_lib.BA_ARR_DC, (p.n_sys,))
    ba2 = BundleAdjuster(b, device=cuda_device, verbose=False)
    motion, _ = ba2.compute_update(10.0)
    assert relerr(-m1.reshape(-1, 6), motion) < 1e-9
    with pytest.raises(_lib.BAError):
        p.set_option(_lib.BA_OPT_SPIN_TIMEOUT_MS, 1e-6)
        ba.compute_update(10.0)


def test_nonfinite_cost_is_reported(cuda_device):
    """window_slam.py:70 runs the reference under numpy.seterr(all='raise'); here a NaN estimate
    surfaces as BA_ERR_NONFINITE from ba_read_scalars instead of a silent NaN cost."""
    from pysfm_b200 import _lib
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    from pysfm_b200 import synthetic
    b = synthetic.make_scene(6, 50, 4, seed=33)
    b._obs_arrays[2][9, 0] = np.nan          # one measurement is NaN
    ba = BundleAdjuster(b, device=cuda_device, verbose=False)
    ba._push(b)
    ba._problem.cost()
    cost, _, status = ba._problem.read_scalars()
    assert not np.isfinite(cost) and status == _lib.BA_ERR_NONFINITE


def test_handle_on_another_device_leaves_the_current_device_alone():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    from pysfm_b200 import synthetic
    torch.cuda.set_device(0)
    ba = BundleAdjuster(synthetic.make_scene(6, 50, 4, seed=34), device="cuda:1", verbose=False)
    ba.compute_update(1.0)
    assert torch.cuda.current_device() == 0


def _assert_same_packing(bundle, camera_ids, track_ids, opt_cam, opt_trk, device):
    from pysfm_b200 import scene
    host = scene.pack_scene(bundle, camera_ids, track_ids, opt_cam, opt_trk)
    dev = scene.pack_scene_device(bundle, camera_ids, track_ids, opt_cam, opt_trk, device)
    assert dev.n_obs == host.n_obs
    assert np.array_equal(dev.pt_ptr, host.pt_ptr)
    assert np.array_equal(dev.obs_cam, host.obs_cam)
    assert np.array_equal(dev.obs_track, host.obs_track)
    assert np.array_equal(dev.obs_uv, host.obs_uv)
    assert np.array_equal(dev.cam_slot, host.cam_slot) and np.array_equal(dev.pt_slot, host.pt_slot)


def test_device_packer_matches_host_packer(cuda_device):
    """ba_pack_observations (count / scan / scatter / per-track sort on the device) against the numpy
    packer: whole bundles, camera / track subsets in shuffled order, fixed cameras in the middle of
    the list, object-backed bundles (Track dicts), empty selections of observations."""
    from pysfm_b200 import synthetic
    rs = np.random.RandomState(3)
    b = synthetic.make_scene(40, 3000, 9, seed=51)
    _assert_same_packing(b, list(range(40)), list(range(3000)), list(range(1, 40)), list(range(3000)), cuda_device)
    cams = [int(c) for c in rs.permutation(40)[:23]]
    trks = [int(t) for t in rs.permutation(3000)[:1700]]
    opt_cam = sorted(int(i) for i in rs.permutation(23)[:15])
    opt_trk = sorted(int(i) for i in rs.permutation(1700)[:900])
    _assert_same_packing(b, cams, trks, opt_cam, opt_trk, cuda_device)
    # the same bundle again: the raw observation list must not be uploaded a second time
    from pysfm_b200 import scene
    import torch
    dev = torch.device(cuda_device)
    ptrs = [t.data_ptr() for t in scene._raw_observations_on_device(b, dev)]
    _assert_same_packing(b, cams[:7], trks[:50], [1, 2, 3], list(range(50)), cuda_device)
    assert [t.data_ptr() for t in scene._raw_observations_on_device(b, dev)] == ptrs
    c = b.clone_params()            # shares the measurements: shares the device copy
    assert [t.data_ptr() for t in scene._raw_observations_on_device(c, dev)] == ptrs
    # object-backed bundle (the reference's fixture: Track objects with measurement dicts)
    from pysfm_b200.bundle import Bundle
    ao = synthetic.make_arrays(6, 40, 6, seed=53)
    table = ao["obs_uv"].reshape(40, 6, 2).transpose(1, 0, 2)          # [camera][track][2]
    mask = rs.rand(6, 40) < 0.8
    mask[0] = True
    bo = Bundle.FromArrays(ao["K"], ao["Rs"], ao["ts"], ao["pts"], table, mask)
    _assert_same_packing(bo, list(range(6)), list(range(40)), list(range(1, 6)), list(range(40)), cuda_device)
    _assert_same_packing(bo, [4, 1, 3], list(range(5, 30)), [0, 2], list(range(0, 25, 2)), cuda_device)
    # a selection whose cameras see nothing of the selected tracks
    a = synthetic.make_arrays(6, 30, 2, seed=52)
    from pysfm_b200.bundle import Bundle as B2
    b2 = B2.FromObservationArrays(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"])
    seen = set(int(c) for c in a["obs_cam"][a["obs_track"] == 0])
    unseen = [c for c in range(6) if c not in seen][:2]
    if len(unseen) == 2:
        _assert_same_packing(b2, unseen, [0], [1], [0], cuda_device)


def test_window_driver_uploads_the_measurements_once(cuda_device):
    """window_slam.run keeps ONE uploaded observation list for all windows: the device
    tensors behind every window's packed scene are views of the same allocation."""
    from pysfm_b200 import window_slam, scene
    g = load_golden("window_slam")
    b = golden_bundle(g)
    ptrs = []

    def hook(i, ba):
        keep = ba._packed.dev["_keep"]
        ptrs.append(tuple(t.data_ptr() for t in keep))
    window_slam.run(b, int(g["win_size"]), device=cuda_device, verbose=False, on_window=hook)
    assert len(ptrs) >= 2 and len(set(ptrs)) == 1


@pytest.mark.parametrize("nc,ranks", [(199, 2), (499, 4), (120, 8)])
def test_distributed_solve_with_virtual_ranks_on_one_gpu(nc, ranks, cuda_device):
    """The distributed reduced solve (ba_solve.cu, chol_dataflow_kernel<true>) needs several GPUs in
    production; its whole protocol -- tile ownership, contribution sums, tile / inverse / y pushes,
    flags, start barrier, replicated backward substitution -- also runs with N "virtual ranks" on ONE
    GPU (N contexts whose peer pointers point at each other, N co-resident launches), which is what
    tools/microbench/dist_solve_bench does: residual against the dense system <= 1e-9, status 0 and
    identical bits of dC on every rank, or it exits non-zero."""
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "tools", "microbench", "dist_solve_bench")
    if not os.path.isfile(exe):
        pytest.skip("tools/microbench/dist_solve_bench not built (__graft_entry__.build() builds it)")
    res = subprocess.run([exe, str(nc), str(ranks), "2"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and res.stdout.rstrip().endswith("OK"), res.stdout[-2000:] + res.stderr[-2000:]
