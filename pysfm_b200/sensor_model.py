"""Measurement-noise models of the reference (``sensor_model.py:7-72``), host side.

On the device the model is two small parameter records (see ``include/ba_b200.h``):
``BA_MODEL_GAUSSIAN`` carries the 2x2 lower-triangular whitening factor ``L`` and
``BA_MODEL_CAUCHY`` carries ``sigma``; the kernels in ``csrc/ba_math.cuh`` evaluate
``residual_from_error`` / ``Jresidual_from_error`` per observation.  The classes here keep
the reference's constructor arguments and method names so user code that builds a
``Bundle`` keeps working; their numpy methods are single-observation conveniences and are
never called by ``BundleAdjuster``.
"""
import numpy as np

MODEL_GAUSSIAN = 0
MODEL_CAUCHY = 1


class GaussianModel(object):
    """``r = L e`` with ``L = chol(inv(cov))``; argument is a covariance (scalar, 2-vector
    or 2x2) exactly as in reference sensor_model.py:8-17."""
    kind = MODEL_GAUSSIAN

    def __init__(self, cov=1.):
        c = np.asarray(cov, dtype=np.float64)
        if c.ndim == 0:
            c = float(c) * np.eye(2)
        elif c.shape == (2,):
            c = np.diag(c)
        assert c.shape == (2, 2)
        self.cov = c
        self.covinv = np.linalg.inv(c)
        self.L = np.linalg.cholesky(self.covinv)

    def device_params(self):
        """(kind, 4 doubles): row-major L."""
        return self.kind, np.ascontiguousarray(self.L, dtype=np.float64).reshape(4)

    def cost_from_error(self, e):
        e = np.asarray(e, dtype=np.float64)
        return float(e.dot(self.covinv).dot(e))

    def residual_from_error(self, e):
        e = np.asarray(e, dtype=np.float64)
        assert e.shape == (2,)
        return self.L.dot(e)

    def Jresidual_from_error(self, e):
        assert np.shape(e) == (2,)
        return self.L

    def clone(self):
        return GaussianModel(self.cov)


class CauchyModel(object):
    """``r = e * sqrt(log(1 + |e|^2/sigma^2)) / |e|``, linear ``e/sigma`` inside |e|<1e-5
    (reference sensor_model.py:37-69)."""
    kind = MODEL_CAUCHY
    LinearWindowAboutZero = 1e-5

    def __init__(self, sigma):
        self.sigma = float(sigma)
        self.sigmasqr = self.sigma * self.sigma

    def device_params(self):
        """(kind, 4 doubles): [sigma, sigma^2, linear window, 0]."""
        return self.kind, np.array([self.sigma, self.sigmasqr, self.LinearWindowAboutZero, 0.0])

    def cost_from_error(self, e):
        e = np.asarray(e, dtype=np.float64)
        return float(np.log(1. + e.dot(e) / self.sigmasqr))

    def residual_from_error(self, e):
        e = np.asarray(e, dtype=np.float64)
        assert e.shape == (2,)
        rho = np.sqrt(e.dot(e))
        if rho < self.LinearWindowAboutZero:
            return e / self.sigma
        return e * (np.sqrt(np.log(1. + rho * rho / self.sigmasqr)) / rho)

    def Jresidual_from_error(self, e):
        e = np.asarray(e, dtype=np.float64)
        assert e.shape == (2,)
        rho = np.sqrt(e.dot(e))
        if rho < self.LinearWindowAboutZero:
            return np.eye(2) / self.sigma
        s = np.sqrt(np.log(1. + rho * rho / self.sigmasqr))
        ee = np.outer(e, e)
        return ee / (rho * s * (rho * rho + self.sigmasqr)) + (rho * np.eye(2) - ee / rho) * (s / (rho * rho))

    def clone(self):
        return CauchyModel(self.sigma)
