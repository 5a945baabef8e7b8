"""Drop-in for the reference's ``bundle_adjuster.BundleAdjuster`` (bundle_adjuster.py:33-343)
whose every numerical stage runs in libba_b200.so on one B200.

Same constructor, ``set_bundle`` selection semantics, ``optimize`` control flow and damping
schedule, staged methods (``prepare_schur_complement`` -> ``apply_damping`` ->
``compute_schur_complement`` -> ``solve_motion_normal_eqns`` -> ``backsubstitute``),
``compute_update`` / ``compute_cost`` / ``update_motion`` / ``update_structure``, attributes
and exception class as the reference.  What differs is where the work happens:

  * ``set_bundle`` packs the selected sub-problem into SoA/CSR arrays and uploads it once;
  * ``optimize`` keeps current and candidate parameters on the device and reads back two
    scalars (cost, candidate cost) + a status word per trial; the accepted bundle is
    materialised on the host once, when the loop ends;
  * ``HCPs`` is materialised densely only on request (it is nc x nt x 6 x 3).

When ``torch.distributed`` is initialised with more than one rank and ``shard=True`` the
points are split across ranks (scene.PackedScene.shard) and the reduced system and the two
costs are all-reduced over NCCL each trial; everything else is rank-local.
"""
from copy import copy

import numpy as np

from . import _lib
from . import scene as _scene
from .bundle import Bundle, Camera


def select(L, mask):
    """Subset of L given either a boolean mask or a list of members of L
    (bundle_adjuster.py:11-24).  Returns (subset, positions of the subset inside L)."""
    L = np.asarray(L)
    mask = np.asarray(mask)
    if mask.dtype.kind == 'b':
        assert len(mask) == len(L)
        subset = L[mask]
    else:
        assert mask.dtype == L.dtype
        assert set(mask.tolist()).issubset(L.tolist()), 'Mask contained some items not in L'
        subset = mask
    where = {v: i for i, v in enumerate(L.tolist())}
    return subset, [where[v] for v in subset.tolist()]


class NormalEquationsIllconditioned(Exception):
    """Raised when the reduced camera system cannot be solved (bundle_adjuster.py:27)."""
    pass


class BundleAdjuster(object):
    # Relative singular-value cutoff for inverting the 3x3 point blocks; None = plain inverse
    # (bundle_adjuster.py:37).
    SCHUR_COMPLIMENT_PINV_THRESHOLD = 1e-5

    def __init__(self, bundle=None, device=None, verbose=True, shard=False):
        self.num_steps = 0
        self.converged = False
        self.costs = []
        self.verbose = verbose
        self._device = device
        self._shard = shard
        self._problem = None
        self._damp_factor = 1.0
        self.trace = []      # one dict per trial step: damping, cost, cand_cost, accepted
        if bundle is not None:
            self.set_bundle(bundle)

    # ------------------------------------------------------------------------------------------
    def _say(self, msg):
        if self.verbose:
            print(msg)

    def _rcond(self):
        t = self.SCHUR_COMPLIMENT_PINV_THRESHOLD
        return -1.0 if t is None else float(t)

    def set_bundle(self, bundle, camera_ids=None, track_ids=None, camera_mask=None, track_mask=None):
        bundle.check_consistency()
        self.bundle = bundle
        ncam, ntrk = len(bundle.cameras), bundle.num_tracks()
        if camera_ids is None:
            self.camera_ids = list(range(ncam))
        else:
            self.camera_ids = [c for c in camera_ids]
            assert isinstance(self.camera_ids[0], (int, np.integer))
            assert min(self.camera_ids) >= 0
            assert max(self.camera_ids) < ncam
            self.camera_ids = [int(c) for c in self.camera_ids]
        if track_ids is None:
            self.track_ids = list(range(ntrk))
        else:
            self.track_ids = [t for t in track_ids]
            assert isinstance(self.track_ids[0], (int, np.integer))
            assert min(self.track_ids) >= 0
            assert max(self.track_ids) < ntrk
            self.track_ids = [int(t) for t in self.track_ids]
        self.camera_id_set = set(self.camera_ids)

        if camera_mask is None:
            assert len(self.camera_ids) > 1, 'Cannot optimize just one camera'
            self.optim_camera_ids = self.camera_ids[1:]          # first camera held fixed
            self.optim_camera_indices = list(range(1, len(self.camera_ids)))
        else:
            ids, idx = select(self.camera_ids, camera_mask)
            self.optim_camera_ids, self.optim_camera_indices = [int(v) for v in ids], idx
        if track_mask is None:
            self.optim_track_ids = copy(self.track_ids)
            self.optim_track_indices = list(range(len(self.track_ids)))
        else:
            ids, idx = select(self.track_ids, track_mask)
            self.optim_track_ids, self.optim_track_indices = [int(v) for v in ids], idx
        assert len(self.optim_track_ids) == len(self.optim_track_indices)
        assert len(self.optim_camera_ids) == len(self.optim_camera_indices)

        self.close()
        self._world, self._rank = 1, 0
        if self._shard:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                self._world, self._rank = dist.get_world_size(), dist.get_rank()
        import torch
        if self._world == 1 and torch.cuda.is_available():
            # the measurements stay on the device across set_bundle calls (uploaded once per bundle);
            # the selection is packed there (ba_pack_observations)
            packed = _scene.pack_scene_device(bundle, self.camera_ids, self.track_ids,
                                              self.optim_camera_indices, self.optim_track_indices, self._device)
        else:
            # sharded: contiguous point ranges are cut from the host image, each rank uploads its own
            packed = _scene.pack_scene(bundle, self.camera_ids, self.track_ids,
                                       self.optim_camera_indices, self.optim_track_indices)
        self._packed_full = packed
        self._packed = packed.shard(self._rank, self._world)
        if self._world > 1:
            # every rank must own at least one point, and all ranks must find out together (a rank
            # that failed alone would leave the others hanging in the next collective)
            sizes = self._gather_objects(self._packed.n_pt)
            assert min(sizes) >= 1, 'cannot shard %d tracks over %d ranks: some rank would own none (%s)' % (
                packed.n_pt, self._world, sizes)
        self._problem = _scene.DeviceProblem(self._packed, self._device)
        self._scalars_t = self._problem.scalars_tensor() if self._world > 1 else None
        if self._world > 1 and self._peer_comm_wanted():
            import torch.distributed as dist

            def gather(b):
                parts = [None] * self._world
                dist.all_gather_object(parts, b)
                return parts
            self._problem.enable_peer_comm(self._rank, self._world, gather)
            dist.barrier()
        self._damp_factor = 1.0
        self._blocks = {}
        self._say('Configured a bundle adjuster for %d cameras, %d tracks' %
                  (len(self.camera_ids), len(self.track_ids)))

    def close(self):
        """Release the device problem.  Sharded over peer memory this is collective: every rank
        unmaps its peers' buffers, the ranks synchronise, then each frees its own."""
        p = self._problem
        if p is None:
            return
        if getattr(self, "_world", 1) > 1 and p.peer_comm:
            import torch.distributed as dist
            p.disconnect_peers()
            if dist.is_initialized():
                dist.barrier()
        p.close()
        self._problem = None

    # ------------------------------------------------------------------------------------------
    # device <-> host state
    def _push(self, bundle):
        """Upload the selected cameras/points of a host bundle into the device state."""
        R, t = bundle.camera_arrays(self.camera_ids)
        pts = np.asarray(bundle.reconstruction, dtype=np.float64)[self._packed.track_ids]
        self._problem.upload_state(R, t, np.ascontiguousarray(pts))

    def _pull_into(self, bundle, which="state"):
        """Write device parameters back into a host bundle (selected cameras/points only)."""
        R, t, x = self._problem.download(which)
        for pos, cid in enumerate(self.camera_ids):
            bundle.cameras[cid].R = R[pos].copy()
            bundle.cameras[cid].t = t[pos].copy()
        if self._world > 1:
            x = self._gather_points(x)
        bundle.reconstruction = np.array(bundle.reconstruction, dtype=np.float64)
        bundle.reconstruction[self.track_ids] = x
        return bundle

    def _gather_objects(self, obj):
        import torch.distributed as dist
        parts = [None] * self._world
        dist.all_gather_object(parts, obj)
        return parts

    def _gather_points(self, x_local):
        """Rows of every rank's shard, concatenated in rank (= point) order."""
        return np.concatenate(self._gather_objects(np.asarray(x_local)), axis=0)

    # ------------------------------------------------------------------------------------------
    # collectives (only when points are sharded)
    def _peer_comm_wanted(self):
        """Peer-memory collectives (ba_comm.cu) when all ranks sit on one node (<= 8 GPUs); the
        NCCL all-reduce otherwise, or when PYSFM_B200_COLLECTIVE=nccl asks for it."""
        import os
        if os.environ.get("PYSFM_B200_COLLECTIVE", "peer").lower() == "nccl":
            return False
        local = int(os.environ.get("LOCAL_WORLD_SIZE", self._world))
        return self._world <= 8 and local == self._world

    def _allreduce_system(self):
        if self._world > 1:
            if self._problem.peer_comm:
                self._problem.allreduce_system()
            else:
                import torch.distributed as dist
                dist.all_reduce(self._problem.sys)

    def _allreduce_costs(self):
        if self._world > 1:
            if self._problem.peer_comm:
                self._problem.allreduce_costs()
            else:
                import torch.distributed as dist
                dist.all_reduce(self._scalars_t[:2])

    # ------------------------------------------------------------------------------------------
    def _trial(self, damping, cam_param_mask=None):
        """One LM trial at the current device state: linearise, eliminate, solve, back-substitute,
        retract into the candidate buffers and evaluate the candidate cost.
        Returns (cost, candidate cost, status)."""
        p = self._problem
        p.linearize_eliminate(damping, self._rcond(), _lib.BA_WANT_SCHUR)
        if not p.dist_solve:
            self._allreduce_system()    # else: the solve sums the ranks' contributions tile by tile itself
        p.solve(cam_param_mask)
        p.backsub_retract_cost()
        self._allreduce_costs()
        cost, cand, status = p.read_scalars()
        if status == _lib.BA_ERR_TIMEOUT:
            raise _lib.BAError("a device-side wait of the solver / a peer barrier timed out (peer lost or kernel fault)")
        return cost, cand, status

    def _split_param_mask(self, param_mask):
        nc, nt = len(self.optim_camera_ids), len(self.optim_track_ids)
        nparams = Bundle.NumCamParams * nc + Bundle.NumPointParams * nt
        if param_mask is None:
            return None
        param_mask = np.asarray(param_mask)
        assert param_mask.dtype.kind == 'b'
        assert param_mask.shape == (nparams,), \
            'param_mask had shape %s but there are %d parameters' % (str(param_mask.shape), nparams)
        assert np.all(param_mask[nc * 6:]), 'Eliminating point parameters not implemented'
        cam = param_mask[:nc * 6]
        return None if np.all(cam) else cam

    def optimize(self, param_mask=None, max_steps=25, init_damping=10., improvement_threshold=1e-4):
        """Levenberg-Marquardt to convergence; control flow of bundle_adjuster.py:117-162."""
        cam_mask = self._split_param_mask(param_mask)
        p = self._problem
        self._push(self.bundle)
        damping = init_damping
        self.num_steps = 0
        self.converged = False
        self.trace = []
        p.cost()
        self._allreduce_costs()
        cur_cost = p.read_scalars()[0]
        self.costs = [cur_cost]
        moved = False
        while not self.converged and self.num_steps < max_steps:
            self.num_steps += 1
            self._say('Step %d: cost=%f, damping=%f' % (self.num_steps, cur_cost, damping))
            while not self.converged and damping < 1e+8:
                _, next_cost, status = self._trial(damping, cam_mask)
                rec = dict(step=self.num_steps, damping=damping, cost=cur_cost, cand_cost=next_cost,
                           status=status, accepted=False)
                self.trace.append(rec)
                if status == _lib.BA_ERR_ILLCONDITIONED:
                    damping *= 10.
                    self.converged = damping > 1e+8
                    continue
                if next_cost < cur_cost:
                    damping *= .1
                    p.accept()
                    moved = True
                    rec['accepted'] = True
                    self.costs.append(next_cost)
                    self.converged = abs(cur_cost - next_cost) < improvement_threshold
                    cur_cost = next_cost
                    break
                else:
                    damping *= 10.
                    self.converged = damping > 1e+8
        if moved:
            self.bundle = self._pull_into(self.bundle.clone_params())
        if self.converged:
            self._say('Converged after %d steps' % self.num_steps)
        else:
            self._say('Failed to converge after %d steps' % self.num_steps)

    # ------------------------------------------------------------------------------------------
    def compute_cost(self, bundle):
        """Sum of squared residuals over optimised tracks x optimised cameras (:165-171)."""
        self._push(bundle)
        self._problem.cost()
        self._allreduce_costs()
        return self._problem.read_scalars()[0]

    def compute_update(self, damping, param_mask=None):
        """(motion update (nc',6), structure update (nt',3)) = minus the damped Gauss-Newton
        solution at the current bundle (:176-208)."""
        cam_mask = self._split_param_mask(param_mask)
        self._push(self.bundle)
        _, _, status = self._trial(damping, cam_mask)
        if status == _lib.BA_ERR_ILLCONDITIONED:
            raise NormalEquationsIllconditioned
        dC, dP = self._fetch_solution()
        return -dC, -dP

    def _fetch_solution(self):
        p = self._problem
        dC = p.get_array(_lib.BA_ARR_DC, (self._packed.n_opt_cam, 6))
        dP_all = p.get_array(_lib.BA_ARR_DP, (self._packed.n_pt, 3))
        dP = dP_all[self._packed.optim_track_indices]
        if self._world > 1:
            dP = self._gather_points(dP)
        return dC, dP

    # ------------------------------------------------------------------------------------------
    # staged interface used by the reference's unit tests (bundle_adjuster_unittest.py:33-36)
    def prepare_schur_complement(self):
        p = self._problem
        sc = self._packed
        self._push(self.bundle)
        p.linearize_eliminate(0.0, self._rcond(), _lib.BA_WANT_BLOCKS)
        self.HCCs = p.get_array(_lib.BA_ARR_HCC, (sc.n_cam, 6, 6))
        self.HPPs = p.get_array(_lib.BA_ARR_HPP, (sc.n_pt, 3, 3))
        self.bCs = p.get_array(_lib.BA_ARR_BC, (sc.n_cam, 6))
        self.bPs = p.get_array(_lib.BA_ARR_BP, (sc.n_pt, 3))
        self.HCP_blocks = p.get_array(_lib.BA_ARR_HCP, (sc.n_obs, 6, 3))
        if self._world > 1:
            # camera blocks are sums over the ranks' points; point and observation blocks concatenate
            self.HCCs = np.sum(self._gather_objects(self.HCCs), axis=0)
            self.bCs = np.sum(self._gather_objects(self.bCs), axis=0)
            self.HPPs = self._gather_points(self.HPPs)
            self.bPs = self._gather_points(self.bPs)
            self.HCP_blocks = self._gather_points(self.HCP_blocks)
        sc = self._packed_full
        self.HPP_invs = np.empty((sc.n_pt, 3, 3))
        self._damp_factor = 1.0
        self._blocks.pop('HCPs', None)

    @property
    def HCPs(self):
        """Dense (nc, nt, 6, 3) cross blocks like the reference's attribute (:107); built on
        demand from the per-observation blocks."""
        if 'HCPs' not in self._blocks:
            sc = self._packed_full
            nbytes = sc.n_cam * sc.n_pt * 18 * 8
            assert nbytes <= (1 << 30), \
                'dense HCPs would take %.1f GB; use HCP_blocks (one 6x3 block per observation)' % (nbytes / 1e9)
            dense = np.zeros((sc.n_cam, sc.n_pt, 6, 3))
            dense[sc.obs_cam, sc.obs_track] = self.HCP_blocks
            self._blocks['HCPs'] = dense
        return self._blocks['HCPs']

    def apply_damping(self, damping):
        """diag *= (1 + damping) on every camera and point block (:238-242, optimize.py:7-9).
        Like the reference this compounds if called twice without a new prepare."""
        f = 1. + damping
        self._damp_factor *= f
        d6, d3 = np.arange(6), np.arange(3)
        self.HCCs[:, d6, d6] *= f
        self.HPPs[:, d3, d3] *= f

    def compute_schur_complement(self):
        """S (nc',nc',6,6) and b (nc',6) of the reduced camera system (:247-278)."""
        p = self._problem
        sc = self._packed
        self._push(self.bundle)
        p.linearize_eliminate(self._damp_factor - 1.0, self._rcond(), _lib.BA_WANT_SCHUR)
        self._allreduce_system()
        self.HPP_invs = p.get_array(_lib.BA_ARR_HPP_INV, (sc.n_pt, 3, 3))
        if self._world > 1:
            self.HPP_invs = self._gather_points(self.HPP_invs)
        A, b = p.system()
        nc = sc.n_opt_cam
        S = A.reshape(nc, 6, nc, 6).transpose(0, 2, 1, 3).copy()
        return S, b.reshape(nc, 6)

    def solve_motion_normal_eqns(self, S, b, param_mask):
        nc = len(self.optim_camera_ids)
        assert np.shape(S) == (nc, nc, 6, 6)
        assert np.shape(b) == (nc, 6)
        assert np.shape(param_mask) == (nc * 6,), 'shape was ' + str(np.shape(param_mask))
        p = self._problem
        A = np.asarray(S, dtype=np.float64).transpose(0, 2, 1, 3).reshape(nc * 6, nc * 6)
        p.upload_system(A, np.asarray(b, dtype=np.float64).reshape(-1))
        mask = np.asarray(param_mask, dtype=bool)
        p.solve(None if mask.all() else mask)
        _, _, status = p.read_scalars()
        if status == _lib.BA_ERR_ILLCONDITIONED:
            raise NormalEquationsIllconditioned
        return p.get_array(_lib.BA_ARR_DC, (nc, 6))

    def backsubstitute(self, dC):
        """dP_i = V_i^-1 (bP_i - sum_j W_ji^T dC_j) for the optimised tracks (:316-331)."""
        p = self._problem
        p.set_solution(np.asarray(dC, dtype=np.float64).reshape(-1, 6))
        p.backsub_retract_cost()
        dP_all = p.get_array(_lib.BA_ARR_DP, (self._packed.n_pt, 3))
        dP = dP_all[self._packed.optim_track_indices]
        return self._gather_points(dP) if self._world > 1 else dP

    # ------------------------------------------------------------------------------------------
    def update_motion(self, delta, bundle):
        """R <- R exp(delta[:3]), t <- t + delta[3:] on the optimised cameras (:334-337)."""
        assert np.shape(delta) == (len(self.optim_camera_ids), 6)
        self._push(bundle)
        self._problem.retract(delta_cam=delta, delta_pt=None)
        R, t, _ = self._problem.download("candidate")
        for idx in self.optim_camera_indices:
            cam = bundle.cameras[self.camera_ids[idx]]
            cam.R, cam.t = R[idx].copy(), t[idx].copy()

    def update_structure(self, delta, bundle):
        """x <- x + delta on the optimised tracks (:340-343)."""
        assert np.shape(delta) == (len(self.optim_track_ids), 3)
        self._push(bundle)
        lo = self._packed.shard_opt_lo      # sharded: this rank retracts its own rows of delta
        local = np.asarray(delta, dtype=np.float64)[lo:lo + self._packed.n_opt_pt]
        self._problem.retract(delta_cam=None, delta_pt=local)
        _, _, x = self._problem.download("candidate")
        x = x[self._packed.optim_track_indices]
        if self._world > 1:
            x = self._gather_points(x)
        bundle.reconstruction = np.array(bundle.reconstruction, dtype=np.float64)
        ids = np.asarray(self.optim_track_ids, dtype=np.int64)
        bundle.reconstruction[ids] = x
