"""N > 1 host logic on CPU: two `gloo` ranks shard the points of one scene, each builds the packed
reduced camera system of ITS shard (with the CPU oracle standing in for the device kernels), the
packed buffers and the costs are all-reduced exactly as BundleAdjuster._allreduce_system /
_allreduce_costs do over NCCL, and the result must equal the single-rank system.  Also checks that
concatenating the shards' point updates in rank order reproduces the global update."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem_of(packed, model):
    from oracle import ba_oracle
    return ba_oracle.Problem(packed.K.reshape(3, 3), packed.cam_R.reshape(-1, 3, 3), packed.cam_t, packed.pts,
                             packed.obs_cam, packed.obs_track, packed.obs_uv, model,
                             np.asarray(packed.optim_camera_indices), np.asarray(packed.optim_track_indices))


def _packed_system(P, damping):
    from oracle import ba_oracle
    from pysfm_b200 import scene
    blocks = ba_oracle.prepare(P)
    ba_oracle.apply_damping(blocks, damping)
    S, b, Vinv = ba_oracle.schur(P, blocks)
    nco = S.shape[0]
    A = S.transpose(0, 2, 1, 3).reshape(6 * nco, 6 * nco)
    return scene.pack_system(A, b.reshape(-1)), blocks, Vinv


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from oracle import ba_oracle
    from pysfm_b200 import scene, synthetic
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = synthetic.make_scene(12, 600, 5, seed=21)
        full = scene.pack_scene(b, range(12), range(600), range(1, 12), range(600))
        mine = full.shard(rank, world)
        model = ('gaussian', np.eye(2))
        damping = 3.0
        packed, blocks, Vinv = _packed_system(_problem_of(mine, model), damping)
        t = torch.from_numpy(packed.copy())
        dist.all_reduce(t)                                   # == BundleAdjuster._allreduce_system
        A, rhs = scene.unpack_system(t.numpy(), full.n_opt_cam)
        P_mine = _problem_of(mine, model)
        dC = ba_oracle.solve_motion(A.reshape(11, 6, 11, 6).transpose(0, 2, 1, 3), rhs.reshape(11, 6))
        dP = ba_oracle.backsubstitute(P_mine, blocks, Vinv, dC)
        cost = torch.tensor([ba_oracle.compute_cost(P_mine), 0.0], dtype=torch.float64)
        dist.all_reduce(cost)                                # == BundleAdjuster._allreduce_costs
        parts = [None] * world
        dist.all_gather_object(parts, np.asarray(dP))        # == BundleAdjuster._gather_points
        if rank == 0:
            np.savez(os.path.join(out_dir, "sharded.npz"), A=A, rhs=rhs, dC=dC, dP=np.concatenate(parts, axis=0),
                     cost=cost.numpy()[0])
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_system_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp
    from oracle import ba_oracle
    from pysfm_b200 import scene, synthetic
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "sharded.npz"))
    b = synthetic.make_scene(12, 600, 5, seed=21)
    full = scene.pack_scene(b, range(12), range(600), range(1, 12), range(600))
    P = _problem_of(full, ('gaussian', np.eye(2)))
    packed, blocks, Vinv = _packed_system(P, 3.0)
    A, rhs = scene.unpack_system(packed, full.n_opt_cam)
    rel = lambda x, y: float(np.max(np.abs(x - y)) / np.max(np.abs(y)))
    assert rel(got["A"], A) < 1e-12
    assert rel(got["rhs"], rhs) < 1e-12
    motion, structure = ba_oracle.compute_update(P, 3.0)
    assert rel(-got["dC"], motion) < 1e-9
    assert rel(-got["dP"], structure) < 1e-9
    assert abs(float(got["cost"]) - ba_oracle.compute_cost(P)) < 1e-9 * ba_oracle.compute_cost(P)


def test_pack_unpack_system_roundtrip():
    from pysfm_b200 import scene
    rng = np.random.RandomState(0)
    for nc in (1, 2, 7):
        G = rng.randn(6 * nc, 6 * nc)
        A = G + G.T
        b = rng.randn(6 * nc)
        p = scene.pack_system(A, b)
        assert p.size == nc * (nc + 1) // 2 * 36 + 6 * nc
        A2, b2 = scene.unpack_system(p, nc)
        assert np.array_equal(A2, A) and np.array_equal(b2, b)
        # block (a, b) sits at index a*nc - a(a-1)/2 + (b-a), rows of the block contiguous
        a_, b_ = nc - 1, nc - 1
        blk = a_ * nc - a_ * (a_ - 1) // 2 + (b_ - a_)
        assert np.array_equal(p[blk * 36:blk * 36 + 36].reshape(6, 6), A[6 * a_:6 * a_ + 6, 6 * b_:6 * b_ + 6])
