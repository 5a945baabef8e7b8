#!/usr/bin/env python
"""BASELINE configs 1 and 3 on one B200: ms per LM iteration (compute_update + candidate cost, device
resident, CUDA events, median of 50) next to the oracle's numpy port on the same box's host cores.
The verbatim reference cannot run on the GPU box; oracle/time_reference.py times it (and the same
port) in the authoring container -> profiles/r2_reference_timing.json.

    python tools/config13_timing.py > gpurun_out/config13_timing.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def gpu_ms(ba, damping=10.0, reps=50):
    import torch
    from pysfm_b200 import _lib
    p = ba._problem
    ba._push(ba.bundle)
    ts = []
    for i in range(reps + 5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p.linearize_eliminate(damping, 1e-5, _lib.BA_WANT_SCHUR)
        p.solve(None)
        p.backsub_retract_cost()
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def port_s(P, reps):
    from oracle import ba_oracle
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        ba_oracle.lm_iteration(P, 10.0)
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def main():
    from conftest import load_golden, golden_bundle, golden_problem
    from oracle import ba_oracle
    from pysfm_b200 import synthetic
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    out = {"host_cpus": os.cpu_count(), "unit_gpu": "ms per LM iteration", "unit_port": "s per LM iteration"}
    a = synthetic.make_arrays(**synthetic.CONFIGS["C1"])
    b1 = synthetic.make_config("C1")
    ba1 = BundleAdjuster(b1, device="cuda:0", verbose=False)
    P1 = ba_oracle.Problem(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"],
                           ('gaussian', np.eye(2)), np.arange(1, len(a["Rs"])), np.arange(len(a["pts"])))
    out["config1"] = {"observations": int(len(a["obs_cam"])), "gpu_ms": gpu_ms(ba1), "oracle_port_s": port_s(P1, 10)}
    g = load_golden("oleg_synthetic")
    b3 = golden_bundle(g)
    ba3 = BundleAdjuster(b3, device="cuda:0", verbose=False)
    nc, nt = len(g["Rs"]), len(g["pts"])
    P3 = ba_oracle.Problem(g["K"], g["Rs"], g["ts"], g["pts"], g["obs_cam"].astype(np.int64), g["obs_track"].astype(np.int64),
                           g["obs_uv"].astype(np.float64), ('gaussian', np.eye(2)), np.arange(1, nc), np.arange(nt))
    out["config3"] = {"observations": int(len(g["obs_cam"])), "gpu_ms": gpu_ms(ba3), "oracle_port_s": port_s(P3, 5)}
    for k in ("config1", "config3"):
        c = out[k]
        c["gpu_obs_per_s"] = c["observations"] / (c["gpu_ms"] * 1e-3)
        c["port_obs_per_s"] = c["observations"] / c["oracle_port_s"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
