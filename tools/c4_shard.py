#!/usr/bin/env python
"""One rank's share of BASELINE config 4 on one B200: 2,000 cameras / 125,000 points / 1.25 M
observations (the full 11,994 x 11,994 reduced camera system, 576 MB packed).  Stage times of an
LM iteration (CUDA events) and size-independent checks: the cost decreases over accepted steps and
the candidate cost the device predicted equals compute_cost of the accepted bundle."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pysfm_b200 import _lib, synthetic
from pysfm_b200.bundle import Bundle
from pysfm_b200.bundle_adjuster import BundleAdjuster

n_cam, n_pt, k = 2000, int(os.environ.get("C4_POINTS", 125000)), 10
t0 = time.perf_counter()
a = synthetic.make_arrays(n_cam, n_pt, k, seed=4)
b = Bundle.FromObservationArrays(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"])
ba = BundleAdjuster(b, device="cuda:0", verbose=False)
t_setup = time.perf_counter() - t0
p = ba._problem
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
stage = np.zeros(3)
reps = 3
for it in range(reps + 1):
    ev[0].record(); p.linearize_eliminate(10.0, 1e-5, _lib.BA_WANT_SCHUR)
    ev[1].record(); p.solve(None)
    ev[2].record(); p.backsub_retract_cost()
    ev[3].record(); torch.cuda.synchronize()
    if it:
        stage += [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
stage /= reps
cost, cand, st = p.read_scalars()
ba.optimize(max_steps=3)
ok_mono = all(c1 < c0 for c0, c1 in zip(ba.costs[:-1], ba.costs[1:]))
ok_cost = abs(ba.compute_cost(ba.bundle) - ba.costs[-1]) < 1e-9 * ba.costs[-1]
out = {"workload": "config 4 shard: %d cameras / %d points / %d observations on one B200" % (n_cam, n_pt, len(a["obs_cam"])),
       "reduced_system": "%d x %d, packed %.0f MB" % (p.n_sys, p.n_sys, p.sys_len * 8 / 1e6),
       "setup_s": t_setup, "stage_ms": {"linearize_eliminate": stage[0], "solve": stage[1], "backsub_retract_cost": stage[2]},
       "solve_tflops_fp64": p.n_sys ** 3 / 3.0 / (stage[1] * 1e-3) / 1e12,
       "first_trial": {"cost": cost, "cand_cost": cand, "status": st},
       "optimize_costs": [float(c) for c in ba.costs], "monotone": bool(ok_mono), "cand_cost_equals_compute_cost": bool(ok_cost)}
print(json.dumps(out, indent=1))
assert st == 0 and ok_mono and ok_cost
