"""GPU stress run (not a pytest): reduced solves of 45 sizes x 3 (dense, masked) against numpy, and
run-to-run reproducibility of optimize() on four scenes.   python tools/stress_gpu.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pysfm_b200 import synthetic
from pysfm_b200.bundle_adjuster import BundleAdjuster
worst = 0.0
rs = np.random.RandomState(0)
for nc in list(range(1, 36)) + [53, 64, 75, 107, 128, 149, 171, 213, 256, 300]:
    b = synthetic.make_scene(nc + 1, 30, min(nc + 1, 3), 31)
    ba = BundleAdjuster(b, device="cuda:0", verbose=False)
    n = 6 * nc
    for rep in range(3):
        G = rs.randn(n, n + 3)
        A = G @ G.T / n + np.diag(rs.rand(n) * 2 + 0.2)
        rhs = rs.randn(n)
        S = A.reshape(nc, 6, nc, 6).transpose(0, 2, 1, 3).copy()
        mask = np.ones(n, bool) if rep == 0 else rs.rand(n) > 0.2
        mask[rs.randint(n)] = True
        x = ba.solve_motion_normal_eqns(S, rhs.reshape(nc, 6), mask).flatten()
        ref = np.zeros(n); ref[mask] = np.linalg.solve(A[mask][:, mask], rhs[mask])
        err = np.max(np.abs(x - ref)) / np.max(np.abs(ref))
        worst = max(worst, err)
        assert err < 1e-9, (nc, rep, err)
print("solver stress ok, worst rel err %.2e" % worst)
# repeated optimize on the same adjuster / fresh adjusters, different sizes
for (ncam, npt, k, seed) in [(6, 200, 3, 1), (40, 3000, 7, 2), (120, 8000, 12, 3), (9, 500, 9, 4)]:
    b = synthetic.make_scene(ncam, npt, k, seed)
    ba = BundleAdjuster(b, device="cuda:0", verbose=False)
    ba.optimize(max_steps=6)
    c1 = list(ba.costs)
    ba2 = BundleAdjuster(b, device="cuda:0", verbose=False)
    ba2.optimize(max_steps=6)
    # (the reductions into the packed system are atomic, so two runs agree to roundoff, not bitwise)
    assert len(c1) == len(ba2.costs) and max(abs(x - y) / abs(y) for x, y in zip(c1, ba2.costs)) < 1e-9, "LM trajectory not reproducible"
    assert all(x > y for x, y in zip(c1[:-1], c1[1:]))
print("optimize determinism ok")
