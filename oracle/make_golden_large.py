#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- golden vectors of the CPU oracle for the BASELINE configurations that are
too large to re-run inside a test or a bench (a config-4 oracle step takes minutes and ~20 GB).

    python oracle/make_golden_large.py [c5] [c4]

Writes tests/golden/config5_sweep.npz and tests/golden/config4_step.npz.  The scenes come from
pysfm_b200.synthetic (seeded numpy RandomState: the same arrays on any machine), the numbers
from oracle/ba_oracle.py, which tests/test_oracle_golden.py pins to the unmodified reference at
the sizes the reference can run.  Only scalars, the camera update (6 nc' doubles) and a strided
sample of the point update are stored, so the fixtures stay small.

  config 5 (500 cameras / 200 k points / 2 M observations, seed 5): one LM step from the
           initial estimate at lambda = 1e-4 .. 1e2 (cost, candidate cost, dC, sampled dP, RMSE) and
           the free-running optimize() trace (every trial: lambda, cost, candidate cost, accepted).
  config 4 (2000 cameras / 1 M points / 10 M observations, seed 4): one LM step at lambda = 10.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ba_oracle  # noqa: E402
from pysfm_b200 import synthetic  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SAMPLE_STRIDE = 997


def problem(cfg):
    a = synthetic.make_arrays(cfg["n_cam"], cfg["n_pt"], cfg["k"], cfg["seed"])
    P = ba_oracle.Problem(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"],
                          ('gaussian', np.eye(2)), np.arange(1, cfg["n_cam"]), np.arange(cfg["n_pt"]))
    return a, P


def one_step(P, damping):
    t0 = time.perf_counter()
    motion, structure = ba_oracle.compute_update(P, damping)
    Pn = ba_oracle.apply_update(P, motion, structure)
    out = dict(cand_cost=ba_oracle.compute_cost(Pn), motion=motion, structure_sample=structure[::SAMPLE_STRIDE].copy(),
               structure_sq=float(np.sum(structure * structure)), structure_absmax=float(np.max(np.abs(structure))),
               rmse_cand=ba_oracle.reprojection_rmse(Pn), seconds=time.perf_counter() - t0)
    return out


def config5():
    cfg = synthetic.CONFIGS["C5"]
    a, P = problem(cfg)
    out = dict(n_cam=cfg["n_cam"], n_pt=cfg["n_pt"], k=cfg["k"], seed=cfg["seed"], sample_stride=SAMPLE_STRIDE,
               cost0=ba_oracle.compute_cost(P), rmse0=ba_oracle.reprojection_rmse(P))
    lambdas = [1e-4, 1e-3, 1e-2, 1e-1, 1.0, 1e1, 1e2]
    out["lambdas"] = np.array(lambdas)
    steps = [one_step(P, lam) for lam in lambdas]
    for key in ("cand_cost", "structure_sq", "structure_absmax", "rmse_cand", "seconds"):
        out["sweep_" + key] = np.array([s[key] for s in steps])
    out["sweep_motion"] = np.stack([s["motion"] for s in steps])
    out["sweep_structure_sample"] = np.stack([s["structure_sample"] for s in steps])
    print("config 5 sweep: %.1f s per step" % np.mean(out["sweep_seconds"]), flush=True)
    t0 = time.perf_counter()
    Pf, info = ba_oracle.optimize(P, max_steps=25)
    out["opt_costs"] = np.array(info["costs"])
    out["opt_num_steps"] = info["num_steps"]
    out["opt_converged"] = info["converged"]
    out["opt_trace_damping"] = np.array([t["damping"] for t in info["trace"]])
    out["opt_trace_cost"] = np.array([t["cost"] for t in info["trace"]])
    out["opt_trace_cand_cost"] = np.array([t["cand_cost"] for t in info["trace"]])
    out["opt_trace_accepted"] = np.array([t["accepted"] for t in info["trace"]])
    out["opt_rmse_final"] = ba_oracle.reprojection_rmse(Pf)
    out["opt_pts_sample"] = Pf.x[::SAMPLE_STRIDE].copy()
    out["opt_Rs"] = Pf.R.copy()
    out["opt_ts"] = Pf.t.copy()
    print("config 5 optimize: %d steps, %d trials, %.1f s" % (info["num_steps"], len(info["trace"]),
                                                               time.perf_counter() - t0), flush=True)
    np.savez_compressed(os.path.join(GOLDEN, "config5_sweep.npz"), **out)


def config4():
    cfg = synthetic.CONFIGS["C4"]
    a, P = problem(cfg)
    out = dict(n_cam=cfg["n_cam"], n_pt=cfg["n_pt"], k=cfg["k"], seed=cfg["seed"], sample_stride=SAMPLE_STRIDE,
               damping=10.0, cost0=ba_oracle.compute_cost(P), rmse0=ba_oracle.reprojection_rmse(P))
    s = one_step(P, 10.0)
    out.update(s)
    print("config 4 step: %.1f s, cost %.9g -> %.9g" % (s["seconds"], out["cost0"], s["cand_cost"]), flush=True)
    np.savez_compressed(os.path.join(GOLDEN, "config4_step.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c5", "c4"]
    if "c5" in which:
        config5()
    if "c4" in which:
        config4()
