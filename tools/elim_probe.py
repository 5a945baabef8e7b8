#!/usr/bin/env python
"""Timing probes for linearize_eliminate at BASELINE config 2: full kernel, without issuing the
bulk reductions (flag 16), without phase D altogether (flag 32), without the scalar atomics of the
reduced right-hand side (flag 48).  Results of the probe runs are
wrong by construction; this only answers "where does the time go"."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pysfm_b200 import _lib, synthetic
from pysfm_b200.bundle import Bundle
from pysfm_b200.bundle_adjuster import BundleAdjuster

a = synthetic.make_arrays(200, 50000, 10, 1)
b = Bundle.FromObservationArrays(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"])
ba = BundleAdjuster(b, device="cuda:0", verbose=False)
p = ba._problem
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
if len(sys.argv) > 1 and sys.argv[1] == "cudamalloc":
    # experiment: system buffer in a plain cudaMalloc allocation (as the peer-comm path has it)
    import ctypes
    mine = (ctypes.c_ubyte * 64)()
    p._chk(p.lib.ba_comm_create(p.h, 0, 2, ctypes.cast(mine, ctypes.c_void_p)), "ba_comm_create")
    print("system buffer rebound to a cudaMalloc allocation")
for name, extra in (("full", 0), ("no bulk issue", 16), ("no phase D", 32), ("no rhs atomics", 48), ("phases A+B(V,bP) only", -2), ("full", 0)):
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p.linearize_eliminate(10.0, 1e-5, (_lib.BA_WANT_SCHUR | extra) if extra >= 0 else 0)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%-16s %.1f us (min of 8; median %.1f)" % (name, 1e3 * min(ts), 1e3 * sorted(ts)[4]))
