"""Seeded synthetic bundle-adjustment scenes for the BASELINE.json configurations.

Same family as the reference's ``bundle_unittest.create_test_bundle`` (bundle_unittest.py:21-77)
scaled up: K = bundle_io's hard-coded intrinsics (bundle_io.py:5-7), points in the unit cube,
cameras a few units back looking at them, every point observed by exactly ``k`` cameras chosen
uniformly without replacement, Gaussian pixel noise.  Everything is produced as flat arrays
(array-backed Bundle) so that 10M observations do not become 10M Python objects.

  C1: make_scene(5, 100, 4, seed=0)          400 obs
  C2: make_scene(200, 50_000, 10, seed=1)    500k obs
  C4: make_scene(2000, 1_000_000, 10, seed=4) 10M obs
  C5: make_scene(500, 200_000, 10, seed=5)   2M obs
"""
import numpy as np

from . import lie
from .bundle import Bundle
from .sensor_model import GaussianModel

K_DEFAULT = np.array([[1500., 0., 740.], [0., 1500., 680.], [0., 0., 1.]])

CONFIGS = {
    "C1": dict(n_cam=5, n_pt=100, k=4, seed=0),
    "C2": dict(n_cam=200, n_pt=50_000, k=10, seed=1),
    "C4": dict(n_cam=2000, n_pt=1_000_000, k=10, seed=4),
    "C5": dict(n_cam=500, n_pt=200_000, k=10, seed=5),
}


def _choose_cameras(rng, n_cam, n_pt, k):
    """(n_pt, k) distinct camera ids per point, uniform without replacement."""
    k = min(k, n_cam)
    if k == n_cam:
        return np.tile(np.arange(n_cam), (n_pt, 1))
    out = np.empty((n_pt, k), dtype=np.int64)
    step = 1 << 16
    for lo in range(0, n_pt, step):
        hi = min(n_pt, lo + step)
        keys = rng.random_sample((hi - lo, n_cam))
        out[lo:hi] = np.argpartition(keys, k, axis=1)[:, :k]
    out.sort(axis=1)
    return out


def make_arrays(n_cam, n_pt, k, seed, noise=1.0, init_sigma=0.01, K=None):
    """Ground truth + noisy measurements + perturbed initial estimate, all as arrays."""
    rng = np.random.RandomState(seed)
    K = K_DEFAULT.copy() if K is None else np.asarray(K, dtype=np.float64)
    pts_true = rng.uniform(-1., 1., (n_pt, 3))
    Rs_true = lie.batch_exp(rng.randn(n_cam, 3) * 0.1)
    ts_true = np.array([0., 0., 5.]) + rng.randn(n_cam, 3) * 0.2
    cams = _choose_cameras(rng, n_cam, n_pt, k)
    kk = cams.shape[1]
    obs_cam = cams.reshape(-1)
    obs_trk = np.repeat(np.arange(n_pt, dtype=np.int64), kk)
    y = np.einsum('oij,oj->oi', Rs_true[obs_cam], pts_true[obs_trk]) + ts_true[obs_cam]
    p = y.dot(K.T)
    uv = p[:, :2] / p[:, 2:3] + rng.randn(len(obs_cam), 2) * noise
    # initial estimate: truth perturbed by init_sigma (cameras by exp(N(0,s^2)), N(0,s^2))
    pts0 = pts_true + rng.randn(n_pt, 3) * init_sigma
    dR = lie.batch_exp(rng.randn(n_cam, 3) * init_sigma)
    Rs0 = np.einsum('nij,njk->nik', Rs_true, dR)
    ts0 = ts_true + rng.randn(n_cam, 3) * init_sigma
    return dict(K=K, Rs=Rs0, ts=ts0, pts=pts0, obs_cam=obs_cam, obs_track=obs_trk, obs_uv=uv,
                Rs_true=Rs_true, ts_true=ts_true, pts_true=pts_true)


def make_scene(n_cam, n_pt, k, seed, noise=1.0, init_sigma=0.01, cov=1.0):
    """Array-backed Bundle at the perturbed initial estimate, GaussianModel(cov)."""
    a = make_arrays(n_cam, n_pt, k, seed, noise=noise, init_sigma=init_sigma)
    return Bundle.FromObservationArrays(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"],
                                        a["obs_uv"], sensor_model=GaussianModel(cov))


def make_config(name, **overrides):
    cfg = dict(CONFIGS[name])
    cfg.update(overrides)
    return make_scene(**cfg)
