"""Points sharded over 2 GPUs of one box (run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`;
skipped on a single-GPU box).  Both collectives are exercised: the peer-memory kernels of
ba_comm.cu (default on one node) and the NCCL all-reduce (PYSFM_B200_COLLECTIVE=nccl); either way
the sharded update must equal the CPU oracle's on the whole scene and the LM trajectory must equal
the single-GPU one."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, collective):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LOCAL_WORLD_SIZE"] = str(world)
    os.environ["PYSFM_B200_COLLECTIVE"] = collective
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from pysfm_b200 import synthetic
        from pysfm_b200.bundle_adjuster import BundleAdjuster
        b = synthetic.make_scene(30, 3001, 6, seed=77)
        ba = BundleAdjuster(b, device="cuda:%d" % rank, verbose=False, shard=True)
        assert ba._problem.peer_comm == (collective == "peer")
        motion, structure = ba.compute_update(2.0)
        cost0 = ba.compute_cost(b)
        ba.optimize(max_steps=4)
        if rank == 0:
            np.savez(os.path.join(out_dir, "mgpu_%s.npz" % collective), motion=motion, structure=structure, cost0=cost0,
                     costs=np.array(ba.costs), Rs=ba.bundle.Rs(), pts=ba.bundle.reconstruction)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("collective", ["peer", "nccl"])
def test_two_gpu_sharded_update_and_trajectory(collective, tmp_path, cuda_device):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import ba_oracle
    from pysfm_b200 import synthetic
    from pysfm_b200.bundle_adjuster import BundleAdjuster
    from conftest import relerr
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), collective), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "mgpu_%s.npz" % collective))
    a = synthetic.make_arrays(30, 3001, 6, seed=77)
    P = ba_oracle.Problem(a["K"], a["Rs"], a["ts"], a["pts"], a["obs_cam"], a["obs_track"], a["obs_uv"],
                          ('gaussian', np.eye(2)), np.arange(1, 30), np.arange(3001))
    m2, s2 = ba_oracle.compute_update(P, 2.0)
    assert relerr(got["motion"], m2) < 1e-7
    assert relerr(got["structure"], s2) < 1e-7
    assert abs(float(got["cost0"]) - ba_oracle.compute_cost(P)) < 1e-10 * ba_oracle.compute_cost(P)
    ba = BundleAdjuster(synthetic.make_scene(30, 3001, 6, seed=77), device=cuda_device, verbose=False)
    ba.optimize(max_steps=4)
    assert relerr(got["costs"], np.array(ba.costs)) < 1e-9
    assert relerr(got["pts"], ba.bundle.reconstruction) < 1e-7
