"""Text loader for the reference's pose/track files (reference bundle_io.py:10-27).

``poses``: one camera per line, 12 floats = row-major 3x4 [R|t].
``tracks``: one track per line, integer triples ``camera u v``.
The intrinsics are the reference's hard-coded ones (bundle_io.py:5-7).  The result is an
array-backed Bundle (no per-track Python objects); ``reconstruction`` is zero until
``triangulate_all`` or the caller fills it, exactly like the reference.
"""
import numpy as np

from .bundle import Bundle

width = 1480
height = 1360
K = np.array([1500, 0, width / 2, 0, 1500, height / 2, 0, 0, 1], float).reshape((3, 3))


def load(tracks_path, cameras_path):
    poses = np.loadtxt(cameras_path).reshape(-1, 3, 4)
    obs_cam, obs_trk, obs_uv = [], [], []
    with open(tracks_path) as f:
        for j, line in enumerate(f):
            vals = np.array(line.split(), dtype=np.int64)
            assert len(vals) % 3 == 0, 'Error at line %d:\n %s' % (j, line)
            v = vals.reshape(-1, 3)
            obs_cam.append(v[:, 0])
            obs_trk.append(np.full(len(v), j, dtype=np.int64))
            obs_uv.append(v[:, 1:].astype(np.float64))
    n_trk = len(obs_cam)
    return Bundle.FromObservationArrays(K, poses[:, :, :3], poses[:, :, 3], np.zeros((n_trk, 3)),
                                        np.concatenate(obs_cam), np.concatenate(obs_trk),
                                        np.concatenate(obs_uv))
