"""Points sharded over the GPUs of one box (run with `gpurun --gpus N -- python -m pytest tests/test_multi_gpu.py -m gpu`;
every case is skipped when the box has fewer GPUs than it needs).  Covered:

  * both collectives -- the peer-memory kernels of ba_comm.cu (default on one node) and the NCCL
    all-reduce (PYSFM_B200_COLLECTIVE=nccl) -- at world sizes 2, 4 and 8: the sharded update must
    equal the CPU oracle's on the whole scene and the LM trajectory the single-GPU one;
  * the DISTRIBUTED reduced solve (ba_solve.cu, DIST: tiles owned by ranks, contributions summed
    and L exchanged over peer memory inside one launch) forced on for small systems, against the
    oracle and bit-for-bit across ranks, including a frozen-parameter mask and an indefinite system;
  * the staged API under sharding (compute_schur_complement returns the GLOBAL system,
    solve_motion_normal_eqns solves the system it is given, backsubstitute / update_structure).
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SCENE = dict(n_cam=30, n_pt=3001, k=6, seed=77)
SCENE_DIST = dict(n_cam=120, n_pt=4001, k=8, seed=78)     # 714 camera parameters: 12 tile rows


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _guard(fn):
    """Run a spawned worker and leave its traceback where the parent can show it."""
    import functools
    import traceback

    @functools.wraps(fn)
    def wrapped(rank, world, port, out_dir, *rest):
        try:
            return fn(rank, world, port, out_dir, *rest)
        except BaseException:
            with open(os.path.join(out_dir, "rank%d.err" % rank), "w") as f:
                f.write(traceback.format_exc())
            raise
    return wrapped


def _spawn(fn, world, args):
    """mp.spawn with the workers' own tracebacks attached to a failure."""
    import glob
    import torch.multiprocessing as mp
    out_dir = args[1]
    try:
        mp.spawn(fn, args=(world,) + tuple(args), nprocs=world, join=True)
    except Exception as exc:
        notes = "".join("\n--- %s ---\n%s" % (os.path.basename(p), open(p).read()) for p in sorted(glob.glob(os.path.join(out_dir, "rank*.err"))))
        raise AssertionError("a rank failed: %s%s" % (exc, notes))


def _init(rank, world, port, env):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LOCAL_WORLD_SIZE"] = str(world)
    for k, v in env.items():
        os.environ[k] = v
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    return dist


@_guard
def _worker(rank, world, port, out_dir, collective):
    dist = _init(rank, world, port, {"PYSFM_B200_COLLECTIVE": collective, "PYSFM_B200_DIST_SOLVE_MIN_TILES": "0"})
    try:
        from pysfm_b200 import synthetic
        from pysfm_b200.bundle_adjuster import BundleAdjuster
        b = synthetic.make_scene(**SCENE)
        ba = BundleAdjuster(b, device="cuda:%d" % rank, verbose=False, shard=True)
        assert ba._problem.peer_comm == (collective == "peer")
        assert not ba._problem.dist_solve
        motion, structure = ba.compute_update(2.0)
        cost0 = ba.compute_cost(b)
        cost0_again = ba.compute_cost(b)       # back-to-back cost reductions (two banks of cost slots)
        # staged API: the GLOBAL reduced system on every rank, and a solve of exactly what is passed in
        ba.prepare_schur_complement()
        ba.apply_damping(2.0)
        S, rhs = ba.compute_schur_complement()
        nc = len(ba.optim_camera_ids)
        dC = ba.solve_motion_normal_eqns(S, rhs, np.ones(6 * nc, bool))
        dP = ba.backsubstitute(dC)
        dC_scaled = ba.solve_motion_normal_eqns(S, 2.0 * rhs, np.ones(6 * nc, bool))   # NOT the system the device last built
        moved = b.clone_params()
        ba.update_structure(-dP, moved)
        ba.optimize(max_steps=4)
        if rank == 0:
            np.savez(os.path.join(out_dir, "mgpu.npz"), motion=motion, structure=structure, cost0=cost0,
                     cost0_again=cost0_again, S=S, rhs=rhs, dC=dC, dP=dP, dC_scaled=dC_scaled, HCCs=ba.HCCs, bPs=ba.bPs,
                     moved_pts=moved.reconstruction, costs=np.array(ba.costs), Rs=ba.bundle.Rs(), pts=ba.bundle.reconstruction)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,collective", [(2, "peer"), (2, "nccl"), (4, "peer"), (8, "peer")])
def test_sharded_update_staged_api_and_trajectory(world, collective, tmp_path, cuda_device):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    from oracle import ba_oracle
    from pysfm_b200 import synthetic
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    from conftest import relerr
    _spawn(_worker, world, (_free_port(), str(tmp_path), collective))
    got = np.load(os.path.join(str(tmp_path), "mgpu.npz"))
    a = synthetic.make_arrays(**SCENE)
    nc, nt = SCENE["n_cam"], SCENE["n_pt"]
    P = ba_oracle.Problem(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"],
                          ('gaussian', np.eye(2)), np.arange(1, nc), np.arange(nt))
    m2, s2 = ba_oracle.compute_update(P, 2.0)
    assert relerr(got["motion"], m2) < 1e-7
    assert relerr(got["structure"], s2) < 1e-7
    c0 = ba_oracle.compute_cost(P)
    assert abs(float(got["cost0"]) - c0) < 1e-10 * c0
    assert float(got["cost0_again"]) == float(got["cost0"])
    # staged path against the oracle's stages
    blocks = ba_oracle.prepare(P)
    assert relerr(got["bPs"], blocks["bPs"]) < 1e-9
    ba_oracle.apply_damping(blocks, 2.0)
    assert relerr(got["HCCs"], blocks["HCCs"]) < 1e-9     # (saved after apply_damping)
    S, rhs, Vinv = ba_oracle.schur(P, blocks)
    assert relerr(got["S"], S) < 1e-9
    assert relerr(got["rhs"], rhs) < 1e-9
    assert relerr(got["dC"], -m2) < 1e-7
    assert relerr(got["dP"], -s2) < 1e-7
    assert relerr(got["dC_scaled"], -2.0 * m2) < 1e-7
    assert relerr(got["moved_pts"], a["pts"] + s2) < 1e-9
    ba = BundleAdjuster(synthetic.make_scene(**SCENE), device=cuda_device, verbose=False)
    ba.optimize(max_steps=4)
    assert relerr(got["costs"], np.array(ba.costs)) < 1e-9
    assert relerr(got["pts"], ba.bundle.reconstruction) < 1e-7


@_guard
def _dist_worker(rank, world, port, out_dir):
    dist = _init(rank, world, port, {"PYSFM_B200_COLLECTIVE": "peer", "PYSFM_B200_DIST_SOLVE_MIN_TILES": "1"})
    try:
        from pysfm_b200 import synthetic, _lib
        from pysfm_b200.bundle_adjuster import BundleAdjuster, NormalEquationsIllconditioned
        b = synthetic.make_scene(**SCENE_DIST)
        ba = BundleAdjuster(b, device="cuda:%d" % rank, verbose=False, shard=True)
        assert ba._problem.peer_comm and ba._problem.dist_solve
        motion, structure = ba.compute_update(10.0)
        nc, nt = len(ba.optim_camera_ids), len(ba.optim_track_ids)
        mask = np.ones(6 * nc + 3 * nt, bool)
        mask[[3, 4, 5, 6 * 7 + 1, 6 * (nc - 1) + 2]] = False            # frozen camera parameters
        motion_m, structure_m = ba.compute_update(10.0, mask)
        # every rank contributes -I / world to the system the distributed solve sums: a non-positive
        # pivot on the chain's rank must be reported on EVERY rank
        import torch
        from pysfm_b200 import scene as _scene
        p = ba._problem
        p.linearize_eliminate(10.0, 1e-5, _lib.BA_WANT_SCHUR)
        neg = _scene.pack_system(-np.eye(6 * nc) / world, np.ones(6 * nc))
        p.sys[:p.sys_len].copy_(torch.as_tensor(neg))
        p.solve(None)
        ill = p.read_scalars()[2] == _lib.BA_ERR_ILLCONDITIONED
        ba.optimize(max_steps=3)
        parts = ba._gather_objects(dict(motion=motion, ill=ill, costs=np.array(ba.costs)))
        if rank == 0:
            same = all(np.array_equal(p["motion"], motion) for p in parts)
            np.savez(os.path.join(out_dir, "dist.npz"), motion=motion, structure=structure, motion_m=motion_m,
                     structure_m=structure_m, same_bits=same, ill=all(p["ill"] for p in parts), costs=np.array(ba.costs),
                     launches=ba._problem.launch_count())
    finally:
        dist.destroy_process_group()


@_guard
def _tc_worker(rank, world, port, out_dir):
    # sharded handles of <= 2 ranks prefer "all-reduce + blocked tcgen05 solve on every rank" over the
    # distributed solve (BA_OPT_TC_OVER_DIST_MAX_WORLD); thresholds lowered so that 12 tile rows qualify
    dist = _init(rank, world, port, {"PYSFM_B200_COLLECTIVE": "peer", "PYSFM_B200_DIST_SOLVE_MIN_TILES": "1",
                                     "PYSFM_B200_TC_MIN_TILES": "4", "PYSFM_B200_TC_WINDOW": "2"})
    try:
        from pysfm_b200 import synthetic
        from pysfm_b200.bundle_adjuster import BundleAdjuster
        b = synthetic.make_scene(**SCENE_DIST)
        ba = BundleAdjuster(b, device="cuda:%d" % rank, verbose=False, shard=True)
        p = ba._problem
        assert p.peer_comm and not p.dist_solve and p.tc_solve_active()
        n0 = p.launch_count()
        motion, structure = ba.compute_update(10.0)
        launches = p.launch_count() - n0
        parts = ba._gather_objects(dict(motion=motion))
        if rank == 0:
            np.savez(os.path.join(out_dir, "tc.npz"), motion=motion, structure=structure, launches=launches,
                     same_bits=all(np.array_equal(q["motion"], motion) for q in parts))
    finally:
        dist.destroy_process_group()


def test_blocked_tcgen05_solve_on_sharded_handles(tmp_path, cuda_device):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import ba_oracle
    from pysfm_b200 import synthetic
    from conftest import relerr
    _spawn(_tc_worker, 2, (_free_port(), str(tmp_path)))
    got = np.load(os.path.join(str(tmp_path), "tc.npz"))
    a = synthetic.make_arrays(**SCENE_DIST)
    nc, nt = SCENE_DIST["n_cam"], SCENE_DIST["n_pt"]
    P = ba_oracle.Problem(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"],
                          ('gaussian', np.eye(2)), np.arange(1, nc), np.arange(nt))
    m2, s2 = ba_oracle.compute_update(P, 10.0)
    assert relerr(got["motion"], m2) < 1e-7
    assert relerr(got["structure"], s2) < 1e-7
    assert bool(got["same_bits"]), "ranks disagree on the bits of dC"
    assert int(got["launches"]) > 20          # elimination + all-reduce + the windows of the blocked solve + back-substitution


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_reduced_solve(world, tmp_path, cuda_device):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    from oracle import ba_oracle
    from pysfm_b200 import synthetic
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    from conftest import relerr
    _spawn(_dist_worker, world, (_free_port(), str(tmp_path)))
    got = np.load(os.path.join(str(tmp_path), "dist.npz"))
    a = synthetic.make_arrays(**SCENE_DIST)
    nc, nt = SCENE_DIST["n_cam"], SCENE_DIST["n_pt"]
    P = ba_oracle.Problem(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"],
                          ('gaussian', np.eye(2)), np.arange(1, nc), np.arange(nt))
    m2, s2 = ba_oracle.compute_update(P, 10.0)
    assert relerr(got["motion"], m2) < 1e-7
    assert relerr(got["structure"], s2) < 1e-7
    cam_mask = np.ones(6 * (nc - 1), bool)
    cam_mask[[3, 4, 5, 6 * 7 + 1, 6 * (nc - 2) + 2]] = False
    m3, s3 = ba_oracle.compute_update(P, 10.0, cam_mask)
    assert relerr(got["motion_m"], m3) < 1e-7
    assert relerr(got["structure_m"], s3) < 1e-7
    assert bool(got["same_bits"]), "ranks disagree on the bits of dC"
    assert bool(got["ill"]), "a non-positive pivot must be reported on every rank"
    ba = BundleAdjuster(synthetic.make_scene(**SCENE_DIST), device=cuda_device, verbose=False)
    ba.optimize(max_steps=3)
    assert relerr(got["costs"], np.array(ba.costs)) < 1e-9
