"""Host-side handle on the tcgen05 trailing update of the blocked reduced solve.

The reference solves the reduced camera system with ``numpy.linalg.solve``
(bundle_adjuster.py:281-312).  For large systems ``ba_solve`` factors it panel by panel and applies
each panel to the trailing matrix as INT8 tensor-core products (``csrc/ba_solve_tc.cuh``); this
module exposes that one step on host arrays (``ba_tc_trailing_update_host``) with every integer
intermediate, which is how the tests hold the device arithmetic to its numpy restatement bit for
bit.  Nothing here is on the product path of ``BundleAdjuster``; the options that steer the path
are ``BA_OPT_TC_*`` / ``PYSFM_B200_TC_*`` (``scene.DeviceProblem``).
"""
import ctypes

import numpy as np

from . import _lib


def trailing_update(A, rhs, window, slices=6, bk=64, saved_rhs=None, device=0, want_intermediates=True):
    """A (ld, ld) float64 whose first K = 64*window columns hold a factored panel L: returns the
    updated copy ``A[i, j] -= sum_k L[i, k] L[j, k]`` for K <= j <= i (lower triangle only), the
    updated right-hand side ``rhs[i] -= sum_k L[i, k] rhs[k]`` (i >= K), and, on request, the INT8
    digit planes (slices, ld_pad, K), the power-of-two row scales (ld_pad,) and the INT32 level
    sums (slices, ld_pad, ld) the device formed on the way."""
    lib = _lib.load()
    A = np.asarray(A, dtype=np.float64)
    ld = A.shape[0]
    assert A.shape == (ld, ld) and ld % 64 == 0
    K = 64 * int(window)
    ld_pad = (ld + 127) // 128 * 128
    Af = np.asfortranarray(A).copy(order='F')          # column-major: element (i, j) at j*ld + i
    r = np.ascontiguousarray(rhs, dtype=np.float64).copy()
    assert r.shape == (ld,)
    sv = None if saved_rhs is None else np.ascontiguousarray(saved_rhs, dtype=np.float64)
    assert sv is None or sv.shape == (64,)
    digits = scale = sums = None
    if want_intermediates:
        digits = np.zeros((slices, ld_pad, K), dtype=np.int8)
        scale = np.zeros(ld_pad, dtype=np.float64)
        sums = np.zeros((slices, ld_pad, ld), dtype=np.int32)

    def vp(x):
        return ctypes.c_void_p(None if x is None else x.ctypes.data)

    rc = lib.ba_tc_trailing_update_host(int(device), ld, int(window), int(slices), int(bk), vp(Af), vp(r), vp(sv),
                                        vp(digits), vp(scale), vp(sums))
    if rc != _lib.BA_OK:
        raise _lib.BAError("ba_tc_trailing_update_host failed with status %d" % rc)
    return dict(A=np.array(Af, order='C'), rhs=r, digits=digits, scale=scale, level_sums=sums)
