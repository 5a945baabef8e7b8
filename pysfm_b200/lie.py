"""SO(3) helpers used on the HOST side of the bundle-adjustment path.

Mirrors the two entry points of the reference that the BA path touches
(``lie.py:21-34`` ``SO3.exp`` and ``lie.py:38-40`` ``SO3.J_expm_x``).  The device
implementation of the same map lives in ``csrc/ba_math.cuh`` (``so3_exp``); this module only
exists so that scene generators, loaders and the ``Camera.perturb`` host API have a rotation
exponential without going through the GPU for a single 3-vector.
"""
import numpy as np

# Below this rotation angle the exponential is the identity (reference lie.py:26-28).
SMALL_ANGLE = 1e-8


def hat(m):
    """3-vector -> 3x3 cross-product matrix (reference ``algebra.skew``, algebra.py:51-56)."""
    a, b, c = (float(v) for v in np.asarray(m, dtype=np.float64).reshape(3))
    out = np.zeros((3, 3))
    out[0, 1], out[0, 2] = -c, b
    out[1, 0], out[1, 2] = c, -a
    out[2, 0], out[2, 1] = -b, a
    return out


class SO3(object):
    @staticmethod
    def exp(m):
        """Rodrigues formula; identity for angles under 1e-8 exactly as the reference does."""
        m = np.asarray(m, dtype=np.float64)
        assert m.shape == (3,), 'shape was ' + str(m.shape)
        theta = float(np.sqrt(m.dot(m)))
        if theta < SMALL_ANGLE:
            return np.eye(3)
        W = hat(m)
        a = np.sin(theta) / theta
        b = (1.0 - np.cos(theta)) / (theta * theta)
        return np.eye(3) + a * W + b * W.dot(W)

    @staticmethod
    def J_expm_x(x):
        """d(exp(m) x)/dm at m=0, i.e. hat(-x)."""
        return hat(-np.asarray(x, dtype=np.float64))


def batch_exp(ms):
    """Vectorised exponential for an (n,3) array -- used by the synthetic scene generator."""
    ms = np.asarray(ms, dtype=np.float64).reshape(-1, 3)
    return np.stack([SO3.exp(m) for m in ms]) if len(ms) else np.zeros((0, 3, 3))
