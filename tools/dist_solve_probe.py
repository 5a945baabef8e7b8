#!/usr/bin/env python
"""Timing + wait-time profile of the DISTRIBUTED reduced solve on the GPUs of one box, without the
35 s it takes to generate BASELINE config 4: every rank gets a share of a random diagonally dominant
packed system of the requested size, the solve runs `reps` times, rank 0 prints the time (CUDA
events, max over ranks) and every rank's profile (ba_solver_profile).

    python -m torch.distributed.run --nproc-per-node 8 tools/dist_solve_probe.py --nc 1999 [--band 2] [--reps 5]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nc", type=int, default=1999)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--band", type=int, default=2)
    ap.add_argument("--strict", type=int, default=0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from pysfm_b200 import _lib, synthetic
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    os.environ["PYSFM_B200_DIST_SOLVE_MIN_TILES"] = "1"
    os.environ["PYSFM_B200_TC_OVER_DIST_MAX_WORLD"] = "0"   # (2 ranks would otherwise all-reduce and run the blocked tcgen05 solve)
    os.environ["PYSFM_B200_DIST_BAND"] = str(args.band)
    os.environ["PYSFM_B200_SOLVER_PROFILE"] = "1"
    os.environ["PYSFM_B200_STRICT_FLAGS"] = str(args.strict)
    b = synthetic.make_scene(args.nc + 1, 8 * world, 2, seed=3)
    ba = BundleAdjuster(b, device=dev, verbose=False, shard=True)
    p = ba._problem
    assert p.dist_solve, "distributed solve not selected"
    nc, n = args.nc, 6 * args.nc
    rng = np.random.RandomState(7)          # the same system on every rank, a 1/world share each
    packed = rng.uniform(-0.01, 0.01, p.sys_len)
    nblk = nc * (nc + 1) // 2
    a = np.arange(nc)
    diag_blk = a * nc - a * (a - 1) // 2
    for r in range(6):
        packed[diag_blk * 36 + r * 7] = 50.0 + r
    packed[nblk * 36:] = rng.uniform(-1, 1, n)
    share = torch.as_tensor(packed / world).to(dev)
    ba._push(b)
    times = []
    clocks = []
    try:
        import pynvml
        pynvml.nvmlInit()
        nv = pynvml.nvmlDeviceGetHandleByIndex(local)
    except Exception:
        nv = None
    p.solver_profile(reset=True)
    for it in range(args.reps + 1):
        p.linearize_eliminate(10.0, 1e-5, _lib.BA_WANT_SCHUR)    # (marks the system as a fresh local contribution)
        p.sys[:p.sys_len].copy_(share)
        dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p.solve(None)
        e1.record()
        if nv is not None:      # SM clock while the solve is running
            import time
            time.sleep(0.002)
            clocks.append(pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM))
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it == 0:
            p.solver_profile(reset=True)     # warm-up launch (task list upload, attribute set)
        else:
            times.append(float(t.item()))
    status = p.read_scalars()[2]
    x = p.get_array(_lib.BA_ARR_DC, (n,))
    prof = p.solver_profile()
    parts = [None] * world
    dist.all_gather_object(parts, dict(rank=rank, status=status, x_hash=float(np.sum(x * np.arange(1, n + 1))), finite=bool(np.all(np.isfinite(x))),
                                       prof=prof, clocks=clocks))
    if rank == 0:
        ok = all(q["status"] == 0 and q["finite"] and q["x_hash"] == parts[0]["x_hash"] for q in parts)
        res = None
        if n <= 6000:    # check the solution on the host
            from pysfm_b200 import scene as _scene
            A, rhs = _scene.unpack_system(packed, nc)
            iu = np.triu_indices(n)
            A[(iu[1], iu[0])] = A[iu]              # the solver reads the upper triangle of every block
            res = float(np.max(np.abs(A.dot(x) - rhs)) / np.max(np.abs(rhs)))
            ok = ok and res < 1e-9
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        out = {"n": n, "tiles": p.ld // 64, "world": world, "band": args.band, "strict": args.strict, "ms_best": min(times), "ms_all": times,
               "fp64_tflops": n ** 3 / 3.0 / (min(times) * 1e-3) / 1e12, "ok": ok, "residual_rel": res, "sm_mhz_during_solve": [q["clocks"] for q in parts],
               "profile_us_per_launch_per_cta": [dict((k, round(v * 1e3 / args.reps / sms, 1)) for k, v in q["prof"].items()) for q in parts]}
        print(json.dumps(out))
    dist.barrier()
    ba.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
