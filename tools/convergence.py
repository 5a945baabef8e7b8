#!/usr/bin/env python
"""BASELINE configs 3 and 5 on one B200: per-iteration cost traces against the CPU side.

  config 3  data/oleg_synthetic (100 cams / 1000 tracks / 100 k obs, via the golden fixture that
            oracle/make_golden.py wrote from the UNMODIFIED reference): free-running optimize(),
            cost per accepted step vs the reference's own trace.
  config 5  500 cams / 200 k points / 2 M obs synthetic: one LM step at lambda = 1e-4 .. 1e2 from the
            same starting point (cost, candidate cost, |dC|, |dP|) vs oracle/ba_oracle.py on the
            host, then a free-running optimize() curve on the GPU with reprojection RMSE, and
            the first `--oracle-steps` accepted steps of the oracle's optimize for comparison.

Writes one JSON document (default profiles/convergence.json).  Test infrastructure: this script is
the only thing outside tests/, smoke() and bench.py that runs the oracle, and only as a checker.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def config3(dev):
    from conftest import load_golden, golden_bundle
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    g = load_golden("oleg_synthetic")
    b = golden_bundle(g)
    ba = BundleAdjuster(b, device=dev, verbose=False)
    steps = int(g["opt_num_steps"])
    t0 = time.perf_counter()
    ba.optimize(max_steps=steps)
    wall = time.perf_counter() - t0
    ref = np.asarray(g["opt_costs"], dtype=np.float64)
    ours = np.asarray(ba.costs, dtype=np.float64)
    return {"workload": "data/oleg_synthetic: 100 cameras / 1000 tracks / 100000 observations, lambda0 = 10",
            "reference": "unmodified pysfm BundleAdjuster.optimize through oracle/refshim.py (tests/golden/oleg_synthetic.npz)",
            "steps": steps, "trials": len(ba.trace), "gpu_wall_s_incl_host_loop": wall,
            "costs_gpu": ours.tolist(), "costs_reference": ref.tolist(),
            "max_rel_cost_diff": rel(ours, ref[:len(ours)]) if len(ours) == len(ref) else None,
            "damping_trace_gpu": [r["damping"] for r in ba.trace],
            "accepted_gpu": [bool(r["accepted"]) for r in ba.trace]}


def config5(dev, oracle_steps, n_pt):
    from oracle import ba_oracle
    from pysfm_b200 import synthetic
    from pysfm_b200.bundle import Bundle
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    cfg = dict(synthetic.CONFIGS["C5"])
    if n_pt:
        cfg["n_pt"] = n_pt
    a = synthetic.make_arrays(**cfg)
    nc, nt = len(a["Rs"]), len(a["pts"])
    b = Bundle.FromObservationArrays(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"])
    P = ba_oracle.Problem(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"],
                          ('gaussian', np.eye(2)), np.arange(1, nc), np.arange(nt))
    ba = BundleAdjuster(b, device=dev, verbose=False)
    p = ba._problem
    sweep = []
    for lam in [1e-4, 1e-3, 1e-2, 1e-1, 1.0, 10.0, 100.0]:
        ba._push(b)
        cost, cand, st = ba._trial(lam)
        dC, dP = ba._fetch_solution()
        t0 = time.perf_counter()
        m2, s2 = ba_oracle.compute_update(P, lam)
        cand2 = ba_oracle.compute_cost(ba_oracle.apply_update(P, m2, s2))
        cpu_s = time.perf_counter() - t0
        sweep.append({"lambda": lam, "cost_gpu": cost, "cand_cost_gpu": cand, "cand_cost_oracle": cand2,
                      "rel_cand_cost": abs(cand - cand2) / abs(cand2), "rel_motion": rel(-dC, m2),
                      "rel_structure": rel(-dP, s2), "norm_motion": float(np.linalg.norm(dC)),
                      "norm_structure": float(np.linalg.norm(dP)), "solve_status": int(st),
                      "oracle_cpu_s": cpu_s})
        print("lambda %g: cand cost gpu %.9e oracle %.9e  rel motion %.2e structure %.2e" % (
            lam, cand, cand2, sweep[-1]["rel_motion"], sweep[-1]["rel_structure"]), file=sys.stderr)
    ba = BundleAdjuster(b, device=dev, verbose=False)
    t0 = time.perf_counter()
    ba.optimize(max_steps=25)
    wall = time.perf_counter() - t0
    R, t, x = ba._problem.download("state")
    rm = ba_oracle.reprojection_rmse(P.with_params(R, t, x))
    out = {"workload": "BASELINE config 5: %d cameras / %d points / %d observations (k=10), sigma = 1 px, camera 0 fixed" % (
               nc, nt, len(a["obs_cam"])),
           "lambda_sweep_single_step": sweep,
           "optimize_gpu": {"costs": [float(c) for c in ba.costs], "num_steps": ba.num_steps, "converged": bool(ba.converged),
                            "trials": len(ba.trace), "wall_s_incl_host_loop": wall,
                            "damping_trace": [r["damping"] for r in ba.trace],
                            "reproj_rmse_px_start": ba_oracle.reprojection_rmse(P), "reproj_rmse_px_end": rm}}
    if oracle_steps > 0:
        t0 = time.perf_counter()
        Pn, info = ba_oracle.optimize(P, max_steps=oracle_steps)
        out["optimize_oracle"] = {"costs": [float(c) for c in info["costs"]], "num_steps": info["num_steps"],
                                  "wall_s": time.perf_counter() - t0,
                                  "reproj_rmse_px_end": ba_oracle.reprojection_rmse(Pn)}
        n = min(len(info["costs"]), len(ba.costs))
        out["optimize_max_rel_cost_diff_first_steps"] = rel(ba.costs[:n], info["costs"][:n])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "convergence.json"))
    ap.add_argument("--oracle-steps", type=int, default=3)
    ap.add_argument("--c5-points", type=int, default=0, help="override the number of points of config 5 (0 = 200000)")
    ap.add_argument("--skip", default="", help="comma list of configs to skip (3,5)")
    args = ap.parse_args()
    import torch
    assert torch.cuda.is_available(), "needs a CUDA device"
    dev = "cuda:0"
    doc = {"gpu": torch.cuda.get_device_name(0)}
    skip = set(args.skip.split(","))
    if "3" not in skip:
        doc["config3"] = config3(dev)
    if "5" not in skip:
        doc["config5"] = config5(dev, args.oracle_steps, args.c5_points)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
