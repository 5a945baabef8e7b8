#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- wall-clock of the UNMODIFIED reference (through oracle/refshim.py) for one
LM iteration = compute_update(lambda) + compute_cost, at the two BASELINE configurations it can run:
config 1 (5 cameras / 100 points / 400 observations) and config 3 (data/oleg_synthetic).  It needs
/root/reference, so it runs in the authoring container only (the GPU box does not have it); the
numpy port of the oracle is timed on the same host beside it, which is the figure `bench.py
--impl reference` reproduces on the GPU box.

    python oracle/time_reference.py [--oleg-iters 2] > profiles/r2_reference_timing.json
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ba_oracle, refshim  # noqa: E402


def timed(fn, iters):
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), ts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--oleg-iters", type=int, default=2)
    args = ap.parse_args()
    refshim.install()
    import bundle as rbundle
    import bundle_adjuster as rba
    import bundle_io as rio
    from pysfm_b200 import synthetic
    out = {"host": {"cpus": os.cpu_count(), "note": "authoring container (no GPU); single process, the reference is pure-Python loops"},
           "unit": "seconds per LM iteration = compute_update(10.0) + compute_cost"}
    sink = io.StringIO()

    # ---- config 1: 5 cameras / 100 points / k = 4, seed 0 (pysfm_b200.synthetic.CONFIGS["C1"]) ----
    a = synthetic.make_arrays(**synthetic.CONFIGS["C1"])
    n_cam, n_pt = len(a["Rs"]), len(a["pts"])
    msm = np.zeros((n_cam, n_pt, 2))
    mask = np.zeros((n_cam, n_pt), bool)
    msm[a["obs_cam"], a["obs_track"]] = a["obs_uv"]
    mask[a["obs_cam"], a["obs_track"]] = True
    b1 = rbundle.Bundle.FromArrays(a["K"], a["Rs"], a["ts"], a["pts"], msm, mask)
    with contextlib.redirect_stdout(sink):
        ba1 = rba.BundleAdjuster(b1)

    def ref_iter_1():
        with contextlib.redirect_stdout(sink):
            ba1.compute_update(10.0)
            ba1.compute_cost(b1)
    P1 = ba_oracle.Problem(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"],
                           ('gaussian', np.eye(2)), np.arange(1, n_cam), np.arange(n_pt))
    ref1, _ = timed(ref_iter_1, 10)
    port1, _ = timed(lambda: ba_oracle.lm_iteration(P1, 10.0), 10)
    out["config1"] = {"cameras": n_cam, "points": n_pt, "observations": int(len(a["obs_cam"])),
                      "reference_verbatim_s": ref1, "oracle_port_s": port1}

    # ---- config 3: data/oleg_synthetic -------------------------------------------------------
    droot = os.path.join(refshim.REFERENCE_ROOT, "data", "oleg_synthetic")
    with contextlib.redirect_stdout(sink):
        b3 = rio.load(os.path.join(droot, "tracks.txt"), os.path.join(droot, "poses.txt"))
        t0 = time.perf_counter()
        b3.triangulate_all()
        tri_s = time.perf_counter() - t0
        ba3 = rba.BundleAdjuster(b3)

    def ref_iter_3():
        with contextlib.redirect_stdout(sink):
            ba3.compute_update(10.0)
            ba3.compute_cost(b3)
    ref3, all3 = timed(ref_iter_3, args.oleg_iters)
    oc, ot, uv = [], [], []
    for j, tr in enumerate(b3.tracks):
        for cid, z in tr.measurements.items():
            oc.append(cid); ot.append(j); uv.append(z)
    Rs = np.array([c.R for c in b3.cameras]); ts = np.array([c.t for c in b3.cameras])
    P3 = ba_oracle.Problem(b3.K, Rs, ts, b3.reconstruction, np.array(oc), np.array(ot), np.array(uv, dtype=float),
                           ('gaussian', np.eye(2)), np.arange(1, len(Rs)), np.arange(len(b3.tracks)))
    port3, _ = timed(lambda: ba_oracle.lm_iteration(P3, 10.0), 5)
    out["config3"] = {"cameras": len(Rs), "points": len(b3.tracks), "observations": len(oc),
                      "reference_verbatim_s": ref3, "reference_verbatim_all_s": all3, "oracle_port_s": port3,
                      "reference_triangulate_all_s": tri_s}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
