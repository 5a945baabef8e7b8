"""CPU-only checks: the C-ABI library builds/loads and exports every symbol the header
declares; the host-side packing / selection / sharding logic; no compute calls (no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden, golden_bundle


def _header_functions():
    text = open(os.path.join(ROOT, "include", "ba_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ba_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from pysfm_b200 import build, _lib
    path = build.build_library()
    assert os.path.isfile(path)
    lib = ctypes.CDLL(path)
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in include/ba_b200.h is not exported" % n
    # the ctypes table binds exactly the header's functions
    assert sorted(_lib.SIGNATURES) == names
    lib.ba_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.ba_version()
    lib.ba_system_ld.restype = ctypes.c_int
    assert lib.ba_system_ld(24) == 64 and lib.ba_system_ld(1194) == 1216 and lib.ba_system_ld(64) == 64
    lib.ba_system_size.restype = ctypes.c_size_t
    assert lib.ba_system_size(199) == 199 * 200 // 2 * 36 + 6 * 199 and lib.ba_system_size(0) == 0


def test_library_is_sm100a_sass():
    import shutil
    import subprocess
    from pysfm_b200 import build
    path = build.build_library()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_hot_kernels_use_the_fp64_tensor_pipe_and_the_tma_path():
    """SASS of the built library: the solver's tile products are DMMA (FP64 tensor pipe), the
    elimination kernel hands its 6x6 blocks to the bulk-reduction engine (UBLKRED), the solver
    publishes tiles with bulk copies (UBLKCP) and stages operands with LDGSTS (cp.async)."""
    import shutil
    import subprocess
    from pysfm_b200 import build
    path = build.build_library()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", path], capture_output=True, text=True).stdout
    per_kernel = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per_kernel[name] = []
        elif name is not None:
            per_kernel[name].append(line)
    def count(kernel_substr, mnemonic):
        return sum(sum(mnemonic in l for l in body) for k, body in per_kernel.items() if kernel_substr in k)
    assert count("chol_dataflow_kernel", "DMMA.8x8x4") > 100
    assert count("chol_dataflow_kernel", "UBLKCP") >= 1
    assert count("chol_dataflow_kernel", "LDGSTS") >= 4
    assert count("linearize_eliminate_kernel", "UBLKRED") >= 4
    # nothing in the product path may fall back to a local-memory stack of any size that matters
    # (ILb0E = the single-GPU solver, ILb1E = the distributed variant with its peer pushes)
    assert count("chol_dataflow_kernelILb0E", "STL") <= 8
    assert count("chol_dataflow_kernelILb1E", "STL") <= 32
    # the distributed variant keeps the bulk copies for the local publication; tiles go to the peers
    # as plain 16-byte stores (STG.E.128), flags with system-scope stores
    assert count("chol_dataflow_kernelILb1E", "UBLKCP") >= 1
    assert count("chol_dataflow_kernelILb1E", "DMMA.8x8x4") > 100
    # large reduced systems: the trailing update of the blocked factorisation is a Blackwell tensor
    # kernel -- tcgen05.mma kind::i8 (UTCIMMA) fed by TMA tensor-map loads (UTMALDG), accumulators read
    # back from tensor memory (LDTM), completion through tcgen05.commit (UTCBAR); product instantiation
    # (DBG = false, 6 slices, 64-byte K steps): 8 merged products per 32 bytes of K, two per stage
    prod = "ozaki_syrk_kernelILi6ELi64ELb0E"
    assert count(prod, "UTCIMMA") == 16
    assert count(prod, "UTMALDG.2D") == 12
    assert count(prod, "LDTM") >= 24
    assert count(prod, "UTCBAR") >= 2
    assert count(prod, "STL") <= 16          # (a handful of spilled words at 168 registers, none in the product loop)
    assert count("ozaki_syrk_kernel", "DMMA") == 0 and count("ozaki_syrk_kernel", "HMMA") == 0


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    from pysfm_b200 import _lib
    b = golden_bundle(load_golden("fixture_gaussian"))
    with pytest.raises(_lib.BAError):
        BundleAdjuster(b, verbose=False)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pysfm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("test oracles", ""), "%s mentions the oracle" % f


def test_select_semantics():
    from pysfm_b200.bundle_adjuster import select
    sub, idx = select([3, 1, 7], [False, True, True])
    assert list(sub) == [1, 7] and idx == [1, 2]
    sub, idx = select([3, 1, 7], np.array([7, 3]))
    assert list(sub) == [7, 3] and idx == [2, 0]
    with pytest.raises(AssertionError):
        select([3, 1, 7], np.array([5]))
    with pytest.raises(AssertionError):
        select([3, 1, 7], [True, False])


def test_pack_scene_layout():
    from pysfm_b200 import scene
    g = load_golden("fixture_cauchy")
    b = golden_bundle(g)
    s = scene.pack_scene(b, [3, 1, 0], [0, 2, 5, 9], [2, 0], [1, 3])
    assert s.n_cam == 3 and s.n_pt == 4 and s.n_opt_cam == 2 and s.n_opt_pt == 2
    assert list(s.cam_slot) == [1, -1, 0]
    assert list(s.pt_slot) == [-1, 0, -1, 1]
    assert s.pt_ptr[0] == 0 and s.pt_ptr[-1] == s.n_obs
    for p in range(s.n_pt):
        seg = slice(s.pt_ptr[p], s.pt_ptr[p + 1])
        slots = s.cam_slot[s.obs_cam[seg]]
        assert np.all(np.diff(slots) >= 0)          # fixed (-1) first, then ascending slot
        assert np.all(s.obs_track[seg] == p)
        tid = s.track_ids[p]
        for o in range(seg.start, seg.stop):
            cid = s.camera_ids[s.obs_cam[o]]
            assert np.array_equal(s.obs_uv[o], b.tracks[tid].measurements[cid])
    # every measurement of a selected track in a selected camera is present exactly once
    want = sum(1 for tid in s.track_ids for cid in b.tracks[tid].measurements if cid in (3, 1, 0))
    assert s.n_obs == want
    assert np.array_equal(s.cam_R[0], b.cameras[3].R.reshape(9))
    assert np.array_equal(s.pts[2], b.reconstruction[5])


def test_object_and_array_backed_bundles_pack_identically():
    from pysfm_b200 import scene
    from pysfm_b200.bundle import Bundle
    g = load_golden("config1_synthetic")
    fast = golden_bundle(g)
    nc, nt = len(g["Rs"]), len(g["pts"])
    msm = np.zeros((nc, nt, 2))
    mask = np.zeros((nc, nt), bool)
    msm[g["obs_cam"], g["obs_track"]] = g["obs_uv"]
    mask[g["obs_cam"], g["obs_track"]] = True
    slow = Bundle.FromArrays(g["K"], g["Rs"], g["ts"], g["pts"], msm, mask)
    a = scene.pack_scene(fast, range(nc), range(nt), range(1, nc), range(nt))
    b = scene.pack_scene(slow, range(nc), range(nt), range(1, nc), range(nt))
    for f in ("pt_ptr", "obs_cam", "obs_uv", "obs_track", "cam_slot", "pt_slot", "cam_R", "cam_t", "pts"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert len(fast.tracks) == nt and fast.tracks[3].measurements.keys() == slow.tracks[3].measurements.keys()


def test_shards_partition_points_and_balance_observations():
    from pysfm_b200 import scene, synthetic
    b = synthetic.make_scene(20, 5000, 6, seed=3)
    full = scene.pack_scene(b, range(20), range(5000), range(1, 20), range(5000))
    for world in (2, 3, 8):
        shards = [full.shard(r, world) for r in range(world)]
        assert sum(s.n_pt for s in shards) == full.n_pt
        assert sum(s.n_obs for s in shards) == full.n_obs
        assert np.array_equal(np.concatenate([s.pts for s in shards]), full.pts)
        assert np.array_equal(np.concatenate([s.obs_uv for s in shards]), full.obs_uv)
        for s in shards:
            assert abs(s.n_obs - full.n_obs / world) <= 6
            assert s.pt_ptr[0] == 0 and s.pt_ptr[-1] == s.n_obs
            assert s.cam_slot is full.cam_slot


def test_bundle_container_api():
    from pysfm_b200.bundle import Bundle, Camera, Track
    b = Bundle(2, 0)
    b.add_track(Track([0, 1], [np.array([1., 2.]), np.array([3., 4.])]))
    assert b.reconstruction.shape == (1, 3)
    with pytest.raises(AssertionError):
        b.check_consistency()          # reconstruction not initialised (bundle.py:154)
    b.reconstruction[0] = [0., 0., 5.]
    b.check_consistency()
    with pytest.raises(Exception):
        b.add_track(Track([5], [np.zeros(2)]))
    c = b.clone_params()
    c.cameras[0].perturb(np.array([0., 0., .1, 1., 0., 0.]))
    c.reconstruction[0, 0] = 9.
    assert np.array_equal(b.cameras[0].R, np.eye(3)) and b.reconstruction[0, 0] == 0.
    assert c.tracks is b.tracks and c.sensor_model is b.sensor_model
    assert b.num_params() == 15
    assert np.allclose(b.predict(0, 0), [0., 0.])
    cam = Camera()
    assert cam.projection_matrix().shape == (3, 4)


def test_sensor_model_device_params():
    from pysfm_b200 import sensor_model
    g = sensor_model.GaussianModel([2., 3.])
    kind, p = g.device_params()
    assert kind == 0 and np.allclose(p, [1 / np.sqrt(2.), 0., 0., 1 / np.sqrt(3.)])
    c = sensor_model.CauchyModel(.4)
    kind, p = c.device_params()
    assert kind == 1 and np.allclose(p[:3], [.4, .16, 1e-5])
    # cost == r.r (reference sensor_model.validate, sensor_model.py:76-99)
    e = np.array([1., 2.])
    for m in (g, c):
        r = m.residual_from_error(e)
        assert abs(m.cost_from_error(e) - r.dot(r)) < 1e-12


def test_so3_exp_host_matches_reference_golden():
    from pysfm_b200 import lie
    g = load_golden("so3_exp")
    for m, R in zip(g["ms"], g["Rs"]):
        assert np.max(np.abs(lie.SO3.exp(m) - R)) < 1e-15


def test_geometry_matches_reference_golden():
    """pysfm_b200.geometry (host side of the sliding-window driver) against geometry.py:5-25."""
    from conftest import load_golden, relerr
    from pysfm_b200 import geometry
    g = load_golden("window_slam")
    R0, t0, R1, t1 = g["Rs"][0], g["ts"][0], g["Rs"][1], g["ts"][1]
    Rr, tr = geometry.relative_pose(R0, t0, R1, t1)
    assert relerr(Rr, g["geo_rel_R"]) < 1e-14 and relerr(tr, g["geo_rel_t"]) < 1e-14
    Rp, tp = geometry.propagate_pose_update(R0, t0, R1, t1, g["Rs"][2], g["ts"][2])
    assert relerr(Rp, g["geo_prop_R"]) < 1e-14 and relerr(tp, g["geo_prop_t"]) < 1e-14
    for f in (geometry.rotation_xy, geometry.rotation_xz, geometry.rotation_yz):
        R = f(0.3)
        assert abs(np.linalg.det(R) - 1.0) < 1e-14 and relerr(R.dot(R.T), np.eye(3)) < 1e-14
