"""Linear multi-view triangulation used to initialise points (reference triangulate.py:6-18).

Initialisation only -- not part of the per-iteration hot path (SURVEY section 8f lists a GPU
version as the first follow-up).  ``algebraic_lsq`` keeps the reference's signature;
``algebraic_lsq_batch`` does all tracks of an array-backed bundle at once.
"""
import numpy as np


def _rows(K, R, t, uv):
    """The two linear constraints (u K2 - K0)(R x + t) = 0, (v K2 - K1)(R x + t) = 0."""
    K = np.asarray(K, dtype=np.float64)
    a0 = K[0] - uv[0] * K[2]
    a1 = K[1] - uv[1] * K[2]
    A = np.vstack((a0.dot(R), a1.dot(R)))
    b = np.array([-(a0.dot(t)), -(a1.dot(t))])
    return A, b


def algebraic_lsq(K, Rs, ts, msms):
    msms = np.asarray(list(msms), dtype=np.float64)
    A = np.empty((2 * len(Rs), 3))
    b = np.empty(2 * len(Rs))
    for i in range(len(Rs)):
        A[2 * i:2 * i + 2], b[2 * i:2 * i + 2] = _rows(K, np.asarray(Rs[i]), np.asarray(ts[i]), msms[i])
    x = np.linalg.lstsq(A, b, rcond=None)[0]
    return x
