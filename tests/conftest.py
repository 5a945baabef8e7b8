import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name):
    path = os.path.join(GOLDEN, name + ".npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def golden_model(g):
    """('gaussian', L) / ('cauchy', sigma) for the oracle, from a golden fixture."""
    from oracle import ba_oracle
    if int(g["model_kind"]) == 0:
        return ('gaussian', ba_oracle.gaussian_L(g["model_param"]))
    return ('cauchy', float(g["model_param"][0, 0]))


def golden_problem(g, prefix):
    """oracle.Problem for the sub-problem a golden fixture's `prefix` stage block selected."""
    from oracle import ba_oracle
    cam_ids = g[prefix + "camera_ids"]
    trk_ids = g[prefix + "track_ids"]
    cam_pos = {int(c): p for p, c in enumerate(cam_ids)}
    trk_pos = {int(t): p for p, t in enumerate(trk_ids)}
    keep = [o for o in range(len(g["obs_cam"]))
            if int(g["obs_cam"][o]) in cam_pos and int(g["obs_track"][o]) in trk_pos]
    oc = np.array([cam_pos[int(g["obs_cam"][o])] for o in keep], dtype=np.int64)
    ot = np.array([trk_pos[int(g["obs_track"][o])] for o in keep], dtype=np.int64)
    uv = g["obs_uv"][keep].astype(np.float64)
    return ba_oracle.Problem(g["K"], g["Rs"][cam_ids], g["ts"][cam_ids], g["pts"][trk_ids], oc, ot, uv,
                             golden_model(g), g[prefix + "optim_camera_indices"],
                             g[prefix + "optim_track_indices"])


def golden_bundle(g):
    """pysfm_b200 Bundle (array-backed) holding a golden fixture's scene."""
    from pysfm_b200.bundle import Bundle
    from pysfm_b200 import sensor_model
    if int(g["model_kind"]) == 0:
        sm = sensor_model.GaussianModel(g["model_param"])
    else:
        sm = sensor_model.CauchyModel(float(g["model_param"][0, 0]))
    return Bundle.FromObservationArrays(g["K"], g["Rs"], g["ts"], g["pts"], g["obs_cam"].astype(np.int64),
                                        g["obs_track"].astype(np.int64), g["obs_uv"].astype(np.float64),
                                        sensor_model=sm)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    denom = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b))) / denom if a.size else 0.0


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return "cuda:0"
