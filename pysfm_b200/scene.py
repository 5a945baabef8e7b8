"""Scene packing (host) and the device-resident problem (torch tensors + ba_handle).

``pack_scene`` turns the reference's object graph (``Bundle.cameras[i].R/.t``,
``Bundle.tracks[j].measurements``, ``Bundle.reconstruction``; bundle.py:54-146) plus the
selection made by ``BundleAdjuster.set_bundle`` (bundle_adjuster.py:54-101) into the SoA /
CSR arrays described in include/ba_b200.h.  ``DeviceProblem`` owns those arrays as torch
CUDA tensors (torch is only the device-memory holder and stream provider) and forwards
every numerical step to libba_b200.so.
"""
import ctypes

import numpy as np

from . import _lib


class PackedScene(object):
    """SoA image of the selected sub-problem.  Cameras, points and slots are host arrays (a few KB
    to a few MB, they change every iteration anyway); the observation arrays -- pt_ptr, obs_cam,
    obs_uv, obs_track -- are host arrays when the host packer built them (pack_scene) and DEVICE
    tensors (`dev`) when the device packer did (pack_scene_device); in that case the host views
    are downloaded on first use (tests, the dense-HCPs property), never on the product path."""
    __slots__ = ("camera_ids", "track_ids", "optim_camera_indices", "optim_track_indices",
                 "K", "model_kind", "model_params", "cam_R", "cam_t", "pts",
                 "_pt_ptr", "_obs_cam", "_obs_uv", "_obs_track", "cam_slot", "pt_slot",
                 "dev", "_n_obs",
                 "shard_lo", "shard_obs_lo", "shard_opt_lo")   # position of a shard inside the whole selection

    def __init__(self):
        self.dev = None
        self._pt_ptr = self._obs_cam = self._obs_uv = self._obs_track = None
        self._n_obs = None

    def _host(self, name):
        v = getattr(self, "_" + name)
        if v is None and self.dev is not None:
            v = self.dev[name].cpu().numpy()
            setattr(self, "_" + name, v)
        return v

    pt_ptr = property(lambda self: self._host("pt_ptr"), lambda self, v: setattr(self, "_pt_ptr", v))
    obs_cam = property(lambda self: self._host("obs_cam"), lambda self, v: setattr(self, "_obs_cam", v))
    obs_uv = property(lambda self: self._host("obs_uv"), lambda self, v: setattr(self, "_obs_uv", v))
    obs_track = property(lambda self: self._host("obs_track"), lambda self, v: setattr(self, "_obs_track", v))

    @property
    def n_cam(self):
        return len(self.camera_ids)

    @property
    def n_pt(self):
        return len(self.track_ids)

    @property
    def n_obs(self):
        return int(self._n_obs) if self._n_obs is not None else int(self.obs_cam.shape[0])

    @property
    def n_opt_cam(self):
        return len(self.optim_camera_indices)

    @property
    def n_opt_pt(self):
        return len(self.optim_track_indices)

    def shard(self, rank, world_size):
        """Contiguous point range for `rank`, balanced by observation count (cameras, K and
        the sensor model are replicated).  Returns a new PackedScene sharing camera arrays."""
        if world_size == 1:
            self.shard_lo = self.shard_obs_lo = self.shard_opt_lo = 0
            return self
        nobs = self.n_obs
        targets = [(nobs * r) // world_size for r in range(world_size + 1)]
        cuts = np.searchsorted(self.pt_ptr, targets, side="left")
        cuts[0], cuts[-1] = 0, self.n_pt
        lo, hi = int(cuts[rank]), int(cuts[rank + 1])
        out = PackedScene()
        for name in ("camera_ids", "optim_camera_indices", "K", "model_kind", "model_params",
                     "cam_R", "cam_t", "cam_slot"):
            setattr(out, name, getattr(self, name))
        out.track_ids = list(self.track_ids[lo:hi])
        o0, o1 = int(self.pt_ptr[lo]), int(self.pt_ptr[hi])
        out.pts = np.ascontiguousarray(self.pts[lo:hi])
        out.pt_ptr = (self.pt_ptr[lo:hi + 1] - o0).astype(np.int32)
        out.obs_cam = np.ascontiguousarray(self.obs_cam[o0:o1])
        out.obs_uv = np.ascontiguousarray(self.obs_uv[o0:o1])
        out.obs_track = np.ascontiguousarray(self.obs_track[o0:o1] - lo)
        # local point slots are renumbered densely; the global slot of local point p is
        # recoverable from optim_track_indices (kept in global numbering)
        local = self.pt_slot[lo:hi]
        out.optim_track_indices = [int(i) for i in np.nonzero(local >= 0)[0]]
        ps = np.full(hi - lo, -1, dtype=np.int32)
        ps[out.optim_track_indices] = np.arange(len(out.optim_track_indices), dtype=np.int32)
        out.pt_slot = ps
        out.shard_lo, out.shard_obs_lo = lo, o0
        out.shard_opt_lo = int(np.count_nonzero(self.pt_slot[:lo] >= 0))   # updated tracks that precede the shard
        return out


def _observation_arrays(bundle, camera_ids, track_ids):
    """(obs_track_pos, obs_cam_pos, obs_uv) for every measurement of a selected track in a
    selected camera -- the set prepare_schur_complement visits (bundle_adjuster.py:224-227)."""
    cam_pos = {cid: p for p, cid in enumerate(camera_ids)}
    fast = getattr(bundle, "_obs_arrays", None)
    if fast is not None:
        o_cam, o_trk, o_uv = fast
        cam_lut = np.full(len(bundle.cameras), -1, dtype=np.int64)
        cam_lut[np.asarray(camera_ids, dtype=np.int64)] = np.arange(len(camera_ids))
        trk_lut = np.full(bundle.num_tracks(), -1, dtype=np.int64)
        trk_lut[np.asarray(track_ids, dtype=np.int64)] = np.arange(len(track_ids))
        cp, tp = cam_lut[o_cam], trk_lut[o_trk]
        keep = (cp >= 0) & (tp >= 0)
        return tp[keep], cp[keep], np.asarray(o_uv, dtype=np.float64)[keep]
    t_list, c_list, uv_list = [], [], []
    for tpos, tid in enumerate(track_ids):
        for cid, z in bundle.tracks[tid].measurements.items():
            p = cam_pos.get(cid)
            if p is not None:
                t_list.append(tpos)
                c_list.append(p)
                uv_list.append(z)
    return (np.asarray(t_list, dtype=np.int64), np.asarray(c_list, dtype=np.int64),
            np.asarray(uv_list, dtype=np.float64).reshape(-1, 2))


def pack_scene(bundle, camera_ids, track_ids, optim_camera_indices, optim_track_indices):
    s = PackedScene()
    s.camera_ids = list(camera_ids)
    s.track_ids = list(track_ids)
    s.optim_camera_indices = [int(i) for i in optim_camera_indices]
    s.optim_track_indices = [int(i) for i in optim_track_indices]
    nc, nt = len(s.camera_ids), len(s.track_ids)
    s.K = np.ascontiguousarray(np.asarray(bundle.K, dtype=np.float64).reshape(9))
    s.model_kind, s.model_params = bundle.sensor_model.device_params()
    s.model_params = np.ascontiguousarray(s.model_params, dtype=np.float64)
    s.cam_R, s.cam_t = bundle.camera_arrays(s.camera_ids)
    s.pts = np.ascontiguousarray(np.asarray(bundle.reconstruction, dtype=np.float64)[s.track_ids])
    s.cam_slot = np.full(nc, -1, dtype=np.int32)
    s.cam_slot[s.optim_camera_indices] = np.arange(len(s.optim_camera_indices), dtype=np.int32)
    s.pt_slot = np.full(nt, -1, dtype=np.int32)
    s.pt_slot[s.optim_track_indices] = np.arange(len(s.optim_track_indices), dtype=np.int32)

    o_trk, o_cam, o_uv = _observation_arrays(bundle, s.camera_ids, s.track_ids)
    # point-major; inside a point: fixed cameras first, then ascending reduced-system slot
    order = np.lexsort((o_cam, s.cam_slot[o_cam], o_trk)) if len(o_trk) else np.zeros(0, np.int64)
    o_trk, o_cam, o_uv = o_trk[order], o_cam[order], o_uv[order]
    counts = np.bincount(o_trk, minlength=nt) if len(o_trk) else np.zeros(nt, np.int64)
    s.pt_ptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int32)
    s.obs_cam = o_cam.astype(np.int32)
    s.obs_track = o_trk.astype(np.int32)
    s.obs_uv = np.ascontiguousarray(o_uv, dtype=np.float64)
    s.shard_lo = s.shard_obs_lo = s.shard_opt_lo = 0
    return s


_RAW_CACHE = {}


def _raw_observations_on_device(bundle, device):
    """(raw_track, raw_cam, raw_uv) of the WHOLE bundle as device tensors, uploaded once per bundle
    and device and shared by every bundle that shares the measurements (clone_params copies the
    parameters, not the tracks: bundle.py:301-310)."""
    import torch
    fast = getattr(bundle, "_obs_arrays", None)
    if fast is not None:
        holder, key_obj = fast, fast
    else:
        holder, key_obj = None, bundle.tracks
    key = (id(key_obj), str(device))
    hit = _RAW_CACHE.get(key)
    if hit is not None and hit[0] is key_obj:
        return hit[1]
    if holder is not None:
        o_cam, o_trk, o_uv = holder
    else:
        t_list, c_list, uv_list = [], [], []
        for j, track in enumerate(bundle.tracks):
            for cid, z in track.measurements.items():
                t_list.append(j)
                c_list.append(cid)
                uv_list.append(z)
        o_trk = np.asarray(t_list, dtype=np.int64)
        o_cam = np.asarray(c_list, dtype=np.int64)
        o_uv = np.asarray(uv_list, dtype=np.float64).reshape(-1, 2)
    n = len(o_cam)
    raw = (torch.as_tensor(np.ascontiguousarray(o_trk, dtype=np.int32)).to(device),
           torch.as_tensor(np.ascontiguousarray(o_cam, dtype=np.int32)).to(device),
           torch.as_tensor(np.ascontiguousarray(o_uv, dtype=np.float64).reshape(-1)).to(device) if n
           else torch.zeros(2, dtype=torch.float64, device=device))
    while len(_RAW_CACHE) >= 4:         # a handful of resident measurement lists at most
        _RAW_CACHE.pop(next(iter(_RAW_CACHE)))
    _RAW_CACHE[key] = (key_obj, raw)    # (holding key_obj keeps its id from being reused)
    return raw


def pack_scene_device(bundle, camera_ids, track_ids, optim_camera_indices, optim_track_indices, device):
    """pack_scene with the observation arrays built ON THE DEVICE (ba_pack_observations): the raw
    measurement list of the bundle is uploaded once (and cached on the bundle), a selection costs
    two look-up tables and four small kernels.  Same layout, same order as pack_scene."""
    import torch
    lib = _lib.load()
    device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    s = PackedScene()
    s.camera_ids = list(camera_ids)
    s.track_ids = list(track_ids)
    s.optim_camera_indices = [int(i) for i in optim_camera_indices]
    s.optim_track_indices = [int(i) for i in optim_track_indices]
    nc, nt = len(s.camera_ids), len(s.track_ids)
    s.K = np.ascontiguousarray(np.asarray(bundle.K, dtype=np.float64).reshape(9))
    s.model_kind, s.model_params = bundle.sensor_model.device_params()
    s.model_params = np.ascontiguousarray(s.model_params, dtype=np.float64)
    s.cam_R, s.cam_t = bundle.camera_arrays(s.camera_ids)
    s.pts = np.ascontiguousarray(np.asarray(bundle.reconstruction, dtype=np.float64)[s.track_ids])
    s.cam_slot = np.full(nc, -1, dtype=np.int32)
    s.cam_slot[s.optim_camera_indices] = np.arange(len(s.optim_camera_indices), dtype=np.int32)
    s.pt_slot = np.full(nt, -1, dtype=np.int32)
    s.pt_slot[s.optim_track_indices] = np.arange(len(s.optim_track_indices), dtype=np.int32)
    s.shard_lo = s.shard_obs_lo = s.shard_opt_lo = 0

    raw_track, raw_cam, raw_uv = _raw_observations_on_device(bundle, device)
    n_raw = int(raw_cam.shape[0])
    cam_lut = np.full(max(len(bundle.cameras), 1), -1, dtype=np.int32)
    cam_lut[np.asarray(s.camera_ids, dtype=np.int64)] = np.arange(nc, dtype=np.int32)
    trk_lut = np.full(max(bundle.num_tracks(), 1), -1, dtype=np.int32)
    trk_lut[np.asarray(s.track_ids, dtype=np.int64)] = np.arange(nt, dtype=np.int32)
    up = lambda a: torch.as_tensor(a).to(device)
    cam_lut_t, trk_lut_t, cam_slot_t = up(cam_lut), up(trk_lut), up(s.cam_slot)
    cap = max(n_raw, 1)
    pt_ptr = torch.empty(nt + 1, dtype=torch.int32, device=device)
    obs_cam = torch.empty(cap, dtype=torch.int32, device=device)
    obs_track = torch.empty(cap, dtype=torch.int32, device=device)
    obs_uv = torch.empty(2 * cap, dtype=torch.float64, device=device)
    scratch = torch.empty(2 * nt, dtype=torch.int32, device=device)
    n_obs = ctypes.c_int(0)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    rc = lib.ba_pack_observations(dev_index, n_raw, vp(raw_track), vp(raw_cam), vp(raw_uv), nt, vp(trk_lut_t), vp(cam_lut_t),
                                  vp(cam_slot_t), vp(pt_ptr), vp(obs_cam), vp(obs_uv), vp(obs_track), vp(scratch),
                                  ctypes.byref(n_obs), stream)
    _lib.check(None, rc, "ba_pack_observations")
    n = int(n_obs.value)
    s._n_obs = n
    s.dev = dict(pt_ptr=pt_ptr, obs_cam=obs_cam[:n], obs_track=obs_track[:n], obs_uv=obs_uv[:2 * n].view(-1, 2),
                 cam_slot=cam_slot_t, _keep=(raw_track, raw_cam, raw_uv))
    return s


def packed_system_index(n_opt_cam):
    """(rows, cols) of the dense reduced-system entry stored at every slot of the packed block
    triangle (include/ba_b200.h): 6x6 blocks (a, b), a <= b, block rows back to back, 36
    row-major doubles each."""
    a, b = np.triu_indices(n_opt_cam)
    rows = (6 * a)[:, None, None] + np.arange(6)[None, :, None] + np.zeros((1, 1, 6), dtype=np.int64)
    cols = (6 * b)[:, None, None] + np.arange(6)[None, None, :] + np.zeros((1, 6, 1), dtype=np.int64)
    return rows.reshape(-1), cols.reshape(-1)


def pack_system(A, b):
    """Dense symmetric A (n, n) and b (n,) -> the packed buffer [upper blocks || b]."""
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    n = b.size
    A = np.asarray(A, dtype=np.float64).reshape(n, n)
    rows, cols = packed_system_index(n // 6)
    return np.concatenate((A[rows, cols], b))


def unpack_system(packed, n_opt_cam):
    """Inverse of pack_system: (A, b) with A mirrored from the stored upper block triangle."""
    n = 6 * n_opt_cam
    packed = np.asarray(packed, dtype=np.float64)
    rows, cols = packed_system_index(n_opt_cam)
    A = np.zeros((n, n))
    A[cols, rows] = packed[:rows.size]
    A[rows, cols] = packed[:rows.size]      # diagonal blocks are stored in full: keep as written
    return A, packed[rows.size:rows.size + n].copy()


def _as_vp(t):
    return ctypes.c_void_p(t.data_ptr())


class DeviceProblem(object):
    """One packed sub-problem resident on one GPU, driven through the C ABI."""

    def __init__(self, scene, device=None):
        import torch
        if not torch.cuda.is_available():
            raise _lib.BAError("pysfm_b200 needs a CUDA device: the bundle-adjustment path has no CPU fallback")
        self.torch = torch
        self.lib = _lib.load()
        self.scene = scene
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        dev = self.device
        self.n_sys = 6 * scene.n_opt_cam
        self.ld = int(self.lib.ba_system_ld(self.n_sys))
        tt = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
        if scene.dev is not None and scene.dev["pt_ptr"].device == dev:
            # built on this device by pack_scene_device: nothing to upload
            self.pt_ptr = scene.dev["pt_ptr"]
            self.obs_cam = scene.dev["obs_cam"] if scene.n_obs else torch.zeros(1, dtype=torch.int32, device=dev)
            self.obs_uv = scene.dev["obs_uv"] if scene.n_obs else torch.zeros(2, dtype=torch.float64, device=dev)
        else:
            self.pt_ptr = tt(scene.pt_ptr, torch.int32)
            self.obs_cam = tt(scene.obs_cam, torch.int32) if scene.n_obs else torch.zeros(1, dtype=torch.int32, device=dev)
            self.obs_uv = tt(scene.obs_uv, torch.float64) if scene.n_obs else torch.zeros(2, dtype=torch.float64, device=dev)
        self.cam_slot = tt(scene.cam_slot, torch.int32)
        self.pt_slot = tt(scene.pt_slot, torch.int32)
        # each parameter set is ONE flat allocation [R | t | x] (views below): a host-driven trial
        # uploads the whole estimate with a single copy (ba_trial_host_packed)
        nR, nt_, nx = 9 * scene.n_cam, 3 * scene.n_cam, 3 * scene.n_pt
        host_flat = np.concatenate([np.asarray(scene.cam_R, dtype=np.float64).reshape(-1),
                                    np.asarray(scene.cam_t, dtype=np.float64).reshape(-1),
                                    np.asarray(scene.pts, dtype=np.float64).reshape(-1)])
        self.params = []
        for _ in range(2):
            flat = torch.as_tensor(host_flat).to(dev)
            self.params.append(dict(flat=flat, R=flat[:nR].view(scene.n_cam, 9), t=flat[nR:nR + nt_].view(scene.n_cam, 3),
                                    x=flat[nR + nt_:].view(scene.n_pt, 3)))
        self.cur = 0
        self.sys_len = int(self.lib.ba_system_size(scene.n_opt_cam))
        self.sys = torch.zeros(max(self.sys_len, 2), dtype=torch.float64, device=dev)
        h = ctypes.c_void_p()
        dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        rc = self.lib.ba_create(dev_index, scene.n_cam, scene.n_pt, scene.n_obs, scene.n_opt_cam,
                                scene.n_opt_pt, ctypes.byref(h))
        _lib.check(None, rc, "ba_create")
        self.h = h
        K = np.ascontiguousarray(scene.K, dtype=np.float64)
        self._chk(self.lib.ba_set_intrinsics(h, K.ctypes.data_as(ctypes.POINTER(ctypes.c_double))), "ba_set_intrinsics")
        mp = np.ascontiguousarray(scene.model_params, dtype=np.float64)
        self._chk(self.lib.ba_set_sensor_model(h, int(scene.model_kind), mp.ctypes.data_as(ctypes.POINTER(ctypes.c_double))),
                  "ba_set_sensor_model")
        self._chk(self.lib.ba_bind_structure(h, _as_vp(self.pt_ptr), _as_vp(self.obs_cam), _as_vp(self.obs_uv),
                                             _as_vp(self.cam_slot), _as_vp(self.pt_slot)), "ba_bind_structure")
        self._bind_params()
        self._chk(self.lib.ba_bind_system(h, _as_vp(self.sys)), "ba_bind_system")
        sp = ctypes.c_void_p()
        self._chk(self.lib.ba_scalars_ptr(h, ctypes.byref(sp)), "ba_scalars_ptr")
        self._scalars_ptr = sp.value
        self.peer_comm = False
        self.dist_solve = False     # sharded + large reduced system: ba_solve is the distributed solve
        import os
        for env, opt in (("PYSFM_B200_SPIN_TIMEOUT_MS", _lib.BA_OPT_SPIN_TIMEOUT_MS),
                         ("PYSFM_B200_STRICT_FLAGS", _lib.BA_OPT_STRICT_FLAGS),
                         ("PYSFM_B200_DIST_SOLVE_MIN_TILES", _lib.BA_OPT_DIST_SOLVE_MIN_TILES),
                         ("PYSFM_B200_DIST_BAND", _lib.BA_OPT_DIST_BAND),
                         ("PYSFM_B200_SOLVER_PROFILE", _lib.BA_OPT_SOLVER_PROFILE),
                         ("PYSFM_B200_TC_MIN_TILES", _lib.BA_OPT_TC_MIN_TILES),
                         ("PYSFM_B200_TC_SLICES", _lib.BA_OPT_TC_SLICES),
                         ("PYSFM_B200_TC_WINDOW", _lib.BA_OPT_TC_WINDOW),
                         ("PYSFM_B200_TC_BK", _lib.BA_OPT_TC_BK),
                         ("PYSFM_B200_TC_OVER_DIST_MAX_WORLD", _lib.BA_OPT_TC_OVER_DIST_MAX_WORLD)):
            if os.environ.get(env):
                self.set_option(opt, float(os.environ[env]))

    # -- plumbing --------------------------------------------------------------------------
    def _chk(self, rc, what):
        _lib.check(self.h, rc, what)

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _bind_params(self):
        s, c = self.params[self.cur], self.params[1 - self.cur]
        self._chk(self.lib.ba_bind_state(self.h, _as_vp(s["R"]), _as_vp(s["t"]), _as_vp(s["x"])), "ba_bind_state")
        self._chk(self.lib.ba_bind_candidate(self.h, _as_vp(c["R"]), _as_vp(c["t"]), _as_vp(c["x"])), "ba_bind_candidate")

    def set_option(self, option, value):
        self._chk(self.lib.ba_set_option(self.h, int(option), float(value)), "ba_set_option")
        self.dist_solve = bool(self.lib.ba_dist_solve_active(self.h))

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.ba_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state -----------------------------------------------------------------------------
    @property
    def state(self):
        return self.params[self.cur]

    @property
    def candidate(self):
        return self.params[1 - self.cur]

    def upload_state(self, cam_R, cam_t, pts, non_blocking=False):
        torch = self.torch
        s = self.state
        s["R"].copy_(torch.as_tensor(cam_R).reshape(s["R"].shape), non_blocking=non_blocking)
        s["t"].copy_(torch.as_tensor(cam_t).reshape(s["t"].shape), non_blocking=non_blocking)
        s["x"].copy_(torch.as_tensor(pts).reshape(s["x"].shape), non_blocking=non_blocking)

    def download(self, which="state"):
        p = self.state if which == "state" else self.candidate
        return (p["R"].cpu().numpy().reshape(-1, 3, 3), p["t"].cpu().numpy().reshape(-1, 3),
                p["x"].cpu().numpy().reshape(-1, 3))

    # -- stages ----------------------------------------------------------------------------
    def linearize_eliminate(self, damping, rcond, flags):
        self._chk(self.lib.ba_linearize_eliminate(self.h, float(damping), float(rcond), int(flags), self._stream()),
                  "ba_linearize_eliminate")

    def solve(self, cam_param_mask=None):
        if cam_param_mask is None:
            mp = None
        else:
            m = np.ascontiguousarray(cam_param_mask, dtype=np.uint8)
            assert m.shape == (self.n_sys,), 'shape was ' + str(m.shape)
            self._mask_keepalive = m
            mp = m.ctypes.data_as(ctypes.c_void_p)
        self._chk(self.lib.ba_solve(self.h, mp, self._stream()), "ba_solve")

    def backsub_retract_cost(self):
        self._chk(self.lib.ba_backsub_retract_cost(self.h, self._stream()), "ba_backsub_retract_cost")

    def cost(self):
        self._chk(self.lib.ba_cost(self.h, self._stream()), "ba_cost")

    def eval_observations(self):
        self._chk(self.lib.ba_eval_observations(self.h, self._stream()), "ba_eval_observations")

    def triangulate(self):
        """Algebraic triangulation of every point from the current cameras (overwrites state points)."""
        self._chk(self.lib.ba_triangulate(self.h, self._stream()), "ba_triangulate")

    def accept(self):
        self._chk(self.lib.ba_accept(self.h), "ba_accept")
        self.cur = 1 - self.cur

    def retract(self, delta_cam=None, delta_pt=None):
        dc = dp = None
        keep = []
        if delta_cam is not None:
            a = np.ascontiguousarray(delta_cam, dtype=np.float64)
            assert a.shape == (self.scene.n_opt_cam, 6), 'shape was ' + str(a.shape)
            keep.append(a)
            dc = a.ctypes.data_as(ctypes.c_void_p)
        if delta_pt is not None:
            a = np.ascontiguousarray(delta_pt, dtype=np.float64)
            assert a.shape == (self.scene.n_opt_pt, 3), 'shape was ' + str(a.shape)
            keep.append(a)
            dp = a.ctypes.data_as(ctypes.c_void_p)
        self._chk(self.lib.ba_retract(self.h, dc, dp, self._stream()), "ba_retract")
        self._chk(self.lib.ba_sync(self.h, self._stream()), "ba_sync")
        del keep

    def set_solution(self, dC):
        a = np.ascontiguousarray(dC, dtype=np.float64)
        assert a.shape == (self.scene.n_opt_cam, 6), 'shape was ' + str(a.shape)
        self._chk(self.lib.ba_set_solution(self.h, a.ctypes.data_as(ctypes.c_void_p), self._stream()), "ba_set_solution")

    def upload_system(self, A, b):
        """Overwrite the bound reduced system with a caller-supplied dense symmetric A and b; the
        next solve factors exactly this system on this rank (ba_upload_system)."""
        n = self.n_sys
        A = np.asarray(A, dtype=np.float64).reshape(n, n)
        host = np.zeros(max(self.sys_len, 2))
        host[:self.sys_len] = pack_system(A, b)
        self._chk(self.lib.ba_upload_system(self.h, host.ctypes.data_as(ctypes.c_void_p), self._stream()),
                  "ba_upload_system")

    def copy_solution_to(self, out_dC, out_dP):
        """D2H of the camera solution and the point update into caller (pinned) torch tensors."""
        assert out_dC.numel() == self.n_sys and out_dP.numel() == 3 * self.scene.n_pt
        self._chk(self.lib.ba_get_array(self.h, _lib.BA_ARR_DC, ctypes.c_void_p(out_dC.data_ptr()), self.n_sys,
                                        self._stream()), "ba_get_array")
        self._chk(self.lib.ba_get_array(self.h, _lib.BA_ARR_DP, ctypes.c_void_p(out_dP.data_ptr()),
                                        3 * self.scene.n_pt, self._stream()), "ba_get_array")

    def trial_host(self, damping, rcond, cam_R=None, cam_t=None, pts=None, out_dC=None, out_dP=None,
                   cam_param_mask=None):
        """One whole LM trial from HOST buffers through ba_trial_host: H2D of the estimate (torch
        CPU tensors, pinned for asynchronous copies; None keeps the device copy), the three
        stages, D2H of dC / dP into caller tensors, one synchronisation.
        Returns (cost, candidate cost, status)."""
        sc = self.scene
        ptr = lambda t, n: None if t is None else (self._host_ptr(t, n))
        mp = None
        if cam_param_mask is not None:
            m = np.ascontiguousarray(cam_param_mask, dtype=np.uint8)
            assert m.shape == (self.n_sys,), 'shape was ' + str(m.shape)
            self._mask_keepalive = m
            mp = m.ctypes.data_as(ctypes.c_void_p)
        cost, cand = ctypes.c_double(), ctypes.c_double()
        status = ctypes.c_int()
        self._chk(self.lib.ba_trial_host(self.h, ptr(cam_R, 9 * sc.n_cam), ptr(cam_t, 3 * sc.n_cam), ptr(pts, 3 * sc.n_pt),
                                         float(damping), float(rcond), mp, ptr(out_dC, self.n_sys),
                                         ptr(out_dP, 3 * sc.n_pt), ctypes.byref(cost), ctypes.byref(cand),
                                         ctypes.byref(status), self._stream()), "ba_trial_host")
        return cost.value, cand.value, status.value

    def trial_host_packed(self, damping, rcond, in_flat, out_flat, cam_param_mask=None):
        """ba_trial_host_packed: `in_flat` = pinned CPU tensor [R | t | x] (or None), `out_flat` =
        pinned CPU tensor of 4 + ld + 3 n_pt doubles.  Returns (cost, cand cost, status, dC view,
        dP view); one H2D, one D2H, one synchronisation."""
        sc = self.scene
        n_in = 12 * sc.n_cam + 3 * sc.n_pt
        n_out = 4 + self.ld + 3 * sc.n_pt
        mp = None
        if cam_param_mask is not None:
            m = np.ascontiguousarray(cam_param_mask, dtype=np.uint8)
            assert m.shape == (self.n_sys,), 'shape was ' + str(m.shape)
            self._mask_keepalive = m
            mp = m.ctypes.data_as(ctypes.c_void_p)
        self._chk(self.lib.ba_trial_host_packed(self.h, None if in_flat is None else self._host_ptr(in_flat, n_in),
                                                float(damping), float(rcond), mp, self._host_ptr(out_flat, n_out),
                                                self._stream()), "ba_trial_host_packed")
        st = float(out_flat[2])
        status = _lib.BA_ERR_TIMEOUT if st == 2.0 else _lib.BA_ERR_ILLCONDITIONED if st != 0.0 else _lib.BA_OK
        return (float(out_flat[0]), float(out_flat[1]), status, out_flat[4:4 + self.n_sys],
                out_flat[4 + self.ld:])

    def _host_ptr(self, t, count):
        assert (not t.is_cuda) and t.dtype == self.torch.float64 and t.is_contiguous() and t.numel() == count, \
            'expected a contiguous float64 CPU tensor of %d elements' % count
        return ctypes.c_void_p(t.data_ptr())

    # -- peer-memory collectives (points sharded over the GPUs of one node) ------------------------
    def enable_peer_comm(self, rank, world, all_gather):
        """Export this rank's comm buffer, exchange the IPC handles with `all_gather(bytes) ->
        [bytes per rank]`, map the peers.  Afterwards `self.sys` views the library-owned buffer."""
        torch = self.torch
        self._torch_sys = self.sys
        mine = (ctypes.c_ubyte * 64)()
        rc = self.lib.ba_comm_create(self.h, int(rank), int(world), ctypes.cast(mine, ctypes.c_void_p))
        handles = all_gather(bytes(mine) if rc == _lib.BA_OK else b"")
        ok = rc == _lib.BA_OK and len(handles) == world and all(len(b) == 64 for b in handles)
        if ok:
            blob = (ctypes.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
            ok = self.lib.ba_comm_connect(self.h, ctypes.cast(blob, ctypes.c_void_p)) == _lib.BA_OK
        # every rank must take the same path: peer memory only if it works everywhere (e.g. not when
        # CUDA_VISIBLE_DEVICES hides the peers from each process)
        if not all(all_gather(b"1" if ok else b"0")[r] == b"1" for r in range(world)):
            self.disable_peer_comm()
            return False
        sp = ctypes.c_void_p()
        self._chk(self.lib.ba_comm_system_ptr(self.h, ctypes.byref(sp)), "ba_comm_system_ptr")

        class _Wrap(object):
            pass
        w = _Wrap()
        w.__cuda_array_interface__ = {"shape": (max(self.sys_len, 2),), "typestr": "<f8", "data": (sp.value, False),
                                      "version": 2}
        self._sys_keepalive = w
        self.sys = torch.as_tensor(w, device=self.device)
        self.peer_comm = True
        self.dist_solve = bool(self.lib.ba_dist_solve_active(self.h))
        return True

    def disconnect_peers(self):
        """First half of the collective tear-down: unmap the peers' buffers (the caller then
        synchronises the ranks before any of them destroys its handle)."""
        if self.peer_comm and getattr(self, "h", None) is not None and self.h.value:
            self._chk(self.lib.ba_comm_disconnect(self.h), "ba_comm_disconnect")

    def disable_peer_comm(self):
        """Back to the caller-owned (torch) system buffer and host-side collectives."""
        if getattr(self, "_torch_sys", None) is not None:
            self.sys = self._torch_sys
        self._chk(self.lib.ba_bind_system(self.h, _as_vp(self.sys)), "ba_bind_system")
        self.peer_comm = False
        self.dist_solve = False

    def allreduce_system(self):
        self._chk(self.lib.ba_allreduce_system(self.h, self._stream()), "ba_allreduce_system")

    def allreduce_costs(self):
        self._chk(self.lib.ba_allreduce_costs(self.h, self._stream()), "ba_allreduce_costs")

    def read_scalars(self):
        cost, cand = ctypes.c_double(), ctypes.c_double()
        status = ctypes.c_int()
        self._chk(self.lib.ba_read_scalars(self.h, ctypes.byref(cost), ctypes.byref(cand), ctypes.byref(status),
                                           self._stream()), "ba_read_scalars")
        return cost.value, cand.value, status.value

    def scalars_tensor(self):
        """torch view of the device scalar record {cost, cand_cost, status, spare} (for the
        cross-rank cost reduction when points are sharded)."""
        torch = self.torch

        class _Wrap(object):
            pass
        w = _Wrap()
        w.__cuda_array_interface__ = {"shape": (4,), "typestr": "<f8", "data": (self._scalars_ptr, False),
                                      "version": 2}
        t = torch.as_tensor(w, device=self.device)
        self._scalars_keepalive = w
        return t

    def get_array(self, which, shape):
        out = np.empty(int(np.prod(shape)), dtype=np.float64)
        self._chk(self.lib.ba_get_array(self.h, int(which), out.ctypes.data_as(ctypes.c_void_p), out.size,
                                        self._stream()), "ba_get_array")
        return out.reshape(shape)

    def system(self):
        """(A, b): A = dense symmetric (n_sys, n_sys) rebuilt from the packed upper blocks."""
        host = np.empty(max(self.sys_len, 1), dtype=np.float64)
        self._chk(self.lib.ba_get_system(self.h, host.ctypes.data_as(ctypes.c_void_p), self.sys_len, self._stream()),
                  "ba_get_system")
        return unpack_system(host[:self.sys_len], self.scene.n_opt_cam)

    def solver_profile(self, reset=True):
        """{slot: milliseconds summed over the solver's CTAs} since the last reset (diagnostics;
        needs set_option(BA_OPT_SOLVER_PROFILE, 1))."""
        out = (ctypes.c_ulonglong * 16)()
        self._chk(self.lib.ba_solver_profile(self.h, ctypes.cast(out, ctypes.c_void_p), int(bool(reset)), self._stream()),
                  "ba_solver_profile")
        return dict((name, out[i] * 1e-6) for i, name in enumerate(_lib.SOLVER_PROFILE_SLOTS))

    def tc_solve_active(self):
        """True when ba_solve takes the blocked path with tcgen05 trailing updates (BA_OPT_TC_MIN_TILES)."""
        return bool(self.lib.ba_tc_solve_active(self.h))

    def tc_solve_profile(self, reset=True):
        """CUDA-event breakdown of the blocked solves run since the last reset with BA_OPT_SOLVER_PROFILE = 1:
        dict of milliseconds PER SOLVE (expand, panels, slices, trailing_updates, backward) + 'solves'."""
        import numpy as np
        out = np.zeros(6)
        self._chk(self.lib.ba_tc_solve_profile(self.h, ctypes.c_void_p(out.ctypes.data), int(bool(reset))), "ba_tc_solve_profile")
        n = max(out[5], 1.0)
        names = ("expand", "panels", "slices", "trailing_updates", "backward")
        d = dict((k, float(out[i] / n)) for i, k in enumerate(names))
        d["solves"] = int(out[5])
        return d

    def launch_count(self):
        return int(self.lib.ba_launch_count(self.h))
