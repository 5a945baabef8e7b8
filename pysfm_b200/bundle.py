"""Host-side scene container with the reference's ``bundle.Bundle`` surface (bundle.py:54-505).

``Camera`` / ``Track`` / ``Bundle`` keep the attribute and method names the reference's
callers use (``cameras[i].R/.t``, ``tracks[j].measurements``, ``reconstruction``, ``K``,
``sensor_model``, ``check_consistency``, ``clone_params``, ``FromArrays`` ...).  They are data
holders: every per-observation quantity (``residual``, ``Jresidual``, ``residuals``,
``Jresiduals``, ``complete_cost``) is evaluated by the CUDA library through
``scene.DeviceProblem`` -- there is no numpy implementation of the measurement model here.

For scenes far larger than an object graph can hold (BASELINE configs 2/4/5) a bundle can be
*array-backed* (``Bundle.FromObservationArrays``): the measurements live in three flat arrays
and ``tracks`` is materialised lazily only if somebody asks for it.
"""
import numpy as np

from . import lie
from . import sensor_model as _sensor_model


def project(K, R, t, x):
    """Pinhole projection pr(K (R x + t)) of one point (reference bundle.py:14-19)."""
    p = np.asarray(K, float).dot(np.asarray(R, float).dot(np.asarray(x, float)) + np.asarray(t, float))
    return p[:2] / p[2]


class Camera(object):
    def __init__(self, R=None, t=None, idx=None):
        self.R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)
        self.t = np.zeros(3) if t is None else np.asarray(t, dtype=np.float64)
        assert self.R.shape == (3, 3)
        assert self.t.shape == (3,)
        self.idx = idx

    @property
    def Rt(self):
        return (self.R, self.t)

    def projection_matrix(self):
        return np.hstack((self.R, self.t.reshape(3, 1)))

    def perturb(self, delta):
        """R <- R exp(delta[:3]), t <- t + delta[3:] (bundle.py:76-80)."""
        delta = np.asarray(delta, dtype=np.float64)
        assert delta.shape == (6,)
        self.R = self.R.dot(lie.SO3.exp(delta[:3]))
        self.t = self.t + delta[3:]
        return self

    def transform(self, R, t):
        self.t = np.dot(R, self.t) + t
        self.R = np.dot(R, self.R)

    def __repr__(self):
        return 'Camera(%s)' % str(self.projection_matrix()).replace('\n', '\n       ')

    __str__ = __repr__


class Track(object):
    """Measurements of one 3-D point, keyed by camera id."""

    def __init__(self, camera_ids=None, measurements=None):
        camera_ids = [] if camera_ids is None else camera_ids
        measurements = [] if measurements is None else measurements
        assert isinstance(camera_ids, list)
        assert len(camera_ids) == len(measurements)
        self.measurements = {int(c): np.asarray(m, dtype=np.float64) for c, m in zip(camera_ids, measurements)}

    def add_measurement(self, camera_id, measurement):
        assert np.shape(measurement) == (2,)
        assert isinstance(camera_id, (int, np.integer))
        self.measurements[int(camera_id)] = np.asarray(measurement, dtype=np.float64)

    def has_measurement(self, camera_id):
        return camera_id in self.measurements

    def get_measurement(self, camera_id):
        return self.measurements[camera_id]

    def camera_ids(self):
        return self.measurements.keys()

    def intersect_camera_ids(self, camera_ids):
        return self.measurements.keys() & set(camera_ids)

    def __repr__(self):
        return 'Track(%s)' % '\n      '.join('%-2d ->  [%10f, %10f]' % (i, m[0], m[1])
                                             for i, m in self.measurements.items())

    __str__ = __repr__


class Bundle(object):
    NumCamParams = 6
    NumPointParams = 3

    def __init__(self, ncameras=0, ntracks=0):
        self.cameras = []
        self._tracks = []
        self._obs_arrays = None      # (obs_cam, obs_track, obs_uv) when array-backed
        self._num_tracks = None
        self.reconstruction = np.zeros((0, 3))
        self.K = np.eye(3)
        self.sensor_model = _sensor_model.GaussianModel(1.)
        for _ in range(ncameras):
            self.add_camera()
        for _ in range(ntracks):
            self.add_track()

    # ---- tracks: list of Track objects, lazily built for array-backed bundles -------------
    @property
    def tracks(self):
        if self._obs_arrays is not None and not self._tracks:
            o_cam, o_trk, o_uv = self._obs_arrays
            built = [Track() for _ in range(self._num_tracks)]
            for c, j, z in zip(o_cam.tolist(), o_trk.tolist(), np.asarray(o_uv)):
                built[j].measurements[c] = z
            self._tracks = built
        return self._tracks

    @tracks.setter
    def tracks(self, value):
        self._tracks = value
        self._obs_arrays = None
        self._num_tracks = None

    def num_tracks(self):
        return self._num_tracks if self._obs_arrays is not None else len(self._tracks)

    # ---- construction ------------------------------------------------------------------------
    def add_camera(self, camera=None):
        if camera is None:
            camera = Camera()
        camera.idx = len(self.cameras)
        self.cameras.append(camera)
        return camera

    def add_track(self, track=None):
        if track is None:
            track = Track()
        else:
            for pos, cid in enumerate(track.camera_ids()):
                if cid < 0 or cid >= len(self.cameras):
                    raise Exception('Invalid camera ID=%d in new track at camera_ids[%d])' % (int(cid), pos))
        assert self._obs_arrays is None, 'add_track on an array-backed bundle'
        self._tracks.append(track)
        self.reconstruction = np.vstack((self.reconstruction, np.zeros(3)))
        return track

    @classmethod
    def FromArrays(cls, K, Rs, ts, pts, measurements, measurement_mask=None):
        """N cameras x M tracks dense measurement table (+ visibility mask), bundle.py:332-364."""
        Rs = np.asarray(Rs, dtype=np.float64)
        ts = np.asarray(ts, dtype=np.float64)
        measurements = np.asarray(measurements, dtype=np.float64)
        if measurement_mask is None:
            measurement_mask = np.ones(measurements.shape[:-1], bool)
        measurement_mask = np.asarray(measurement_mask, dtype=bool)
        assert len(Rs) == len(ts)
        assert Rs.shape[1:] == (3, 3)
        assert ts.shape[1:] == (3,)
        assert measurements.shape[0] == len(Rs)
        assert measurements.shape[2] == 2
        assert measurement_mask.shape == measurements.shape[:-1]
        b = cls()
        b.K = np.array(K, dtype=np.float64)
        for R, t in zip(Rs, ts):
            b.add_camera(Camera(R.copy(), t.copy()))
        for j in range(measurements.shape[1]):
            seen = [int(i) for i in np.nonzero(measurement_mask[:, j])[0]]
            b.add_track(Track(seen, measurements[seen, j]))
        b.reconstruction = np.array(pts, dtype=np.float64)
        return b

    @classmethod
    def FromObservationArrays(cls, K, Rs, ts, pts, obs_cam, obs_track, obs_uv, sensor_model=None):
        """Array-backed bundle: observation o is camera obs_cam[o] seeing track obs_track[o]
        at pixel obs_uv[o].  No per-track Python objects are created."""
        b = cls()
        b.K = np.array(K, dtype=np.float64)
        for R, t in zip(np.asarray(Rs, dtype=np.float64), np.asarray(ts, dtype=np.float64)):
            b.add_camera(Camera(R.copy(), t.copy()))
        b.reconstruction = np.array(pts, dtype=np.float64)
        b._num_tracks = len(b.reconstruction)
        b._obs_arrays = (np.asarray(obs_cam, dtype=np.int64), np.asarray(obs_track, dtype=np.int64),
                         np.asarray(obs_uv, dtype=np.float64).reshape(-1, 2))
        if sensor_model is not None:
            b.sensor_model = sensor_model
        return b

    # ---- consistency (bundle.py:148-166) --------------------------------------------------------
    def check_consistency(self):
        assert self.sensor_model is not None
        assert np.shape(self.K) == (3, 3), 'shape was ' + str(np.shape(self.K))
        assert np.shape(self.reconstruction) == (self.num_tracks(), 3), \
            'shape was ' + str(np.shape(self.reconstruction))
        assert np.sum(np.square(self.reconstruction)) > 1e-8, 'reconstruction must be initialized'
        ncam = len(self.cameras)
        if self._obs_arrays is not None:
            o_cam, o_trk, _ = self._obs_arrays
            assert o_cam.min() >= 0 and o_cam.max() < ncam, 'measurement refers to a missing camera'
            assert np.all(np.bincount(o_trk, minlength=self.num_tracks()) > 0), 'track without measurements'
        else:
            for track in self._tracks:
                ids = list(track.camera_ids())
                assert len(ids) > 0 and min(ids) >= 0 and max(ids) < ncam, \
                    'There are %d cameras but track has a measurements for %s' % (ncam, str(ids))
        for camera in self.cameras:
            assert camera.R.shape == (3, 3)
            assert camera.t.shape == (3,)

    # ---- simple accessors ----------------------------------------------------------------------
    def Rs(self):
        return np.array([cam.R for cam in self.cameras])

    def ts(self):
        return np.array([cam.t for cam in self.cameras])

    def camera_arrays(self, camera_ids=None):
        """(n,9) rotations and (n,3) translations of the selected cameras, C-contiguous."""
        ids = range(len(self.cameras)) if camera_ids is None else camera_ids
        R = np.array([self.cameras[i].R for i in ids], dtype=np.float64).reshape(-1, 9)
        t = np.array([self.cameras[i].t for i in ids], dtype=np.float64).reshape(-1, 3)
        return np.ascontiguousarray(R), np.ascontiguousarray(t)

    def projection_matrices(self):
        return np.array([cam.projection_matrix() for cam in self.cameras])

    def points(self):
        return self.reconstruction

    def measurement(self, i, j):
        return self.tracks[j].get_measurement(i)

    def measurement_ids(self, track_indices=None):
        if track_indices is None:
            track_indices = range(self.num_tracks())
        return ((i, j) for j in track_indices for i in self.tracks[j].camera_ids())

    def measurement_ids_for_cameras(self, cameras_to_include, track_indices=None):
        if track_indices is None:
            track_indices = range(self.num_tracks())
        return ((i, j) for j in track_indices for i in self.tracks[j].intersect_camera_ids(cameras_to_include))

    def num_params(self):
        return len(self.cameras) * Bundle.NumCamParams + self.num_tracks() * Bundle.NumPointParams

    def predict(self, i, j):
        return project(self.K, self.cameras[i].R, self.cameras[i].t, self.reconstruction[j])

    def reproj_error(self, i, j):
        return self.predict(i, j) - self.measurement(i, j)

    # ---- device-evaluated measurement model ---------------------------------------------------
    def _evaluate(self, camera_ids=None, track_ids=None):
        """Run ba_eval_observations over the selected cameras x tracks.  Returns the packed
        scene (for index bookkeeping) and per-observation (r, Jc, Jp) in packed order."""
        from . import scene as _scene
        from . import _lib
        cams = list(range(len(self.cameras))) if camera_ids is None else list(camera_ids)
        trks = list(range(self.num_tracks())) if track_ids is None else list(track_ids)
        packed = _scene.pack_scene(self, cams, trks, range(len(cams)), range(len(trks)))
        prob = _scene.DeviceProblem(packed)
        try:
            prob.eval_observations()
            n = packed.n_obs
            r = prob.get_array(_lib.BA_ARR_RESIDUAL, (n, 2))
            Jc = prob.get_array(_lib.BA_ARR_JC, (n, 2, 6))
            Jp = prob.get_array(_lib.BA_ARR_JP, (n, 2, 3))
        finally:
            prob.close()
        return packed, r, Jc, Jp

    def _reference_order(self, packed, camera_ids, track_ids, by_camera_list):
        """Permutation taking packed observation order to the reference's enumeration order:
        tracks in order, then either the track's own camera order (measurement_ids) or the
        order of `camera_ids` (residuals_partial)."""
        key = {}
        for o in range(packed.n_obs):
            key[(packed.camera_ids[packed.obs_cam[o]], packed.track_ids[packed.obs_track[o]])] = o
        order = []
        for j in track_ids:
            cams = camera_ids if by_camera_list else self.tracks[j].camera_ids()
            for i in cams:
                if (i, j) in key:
                    order.append(key[(i, j)])
        return np.asarray(order, dtype=np.int64)

    def residual(self, i, j):
        _, r, _, _ = self._evaluate([i], [j])
        return r[0]

    def Jresidual(self, i, j):
        _, _, Jc, Jp = self._evaluate([i], [j])
        return Jc[0], Jp[0]

    def residuals(self):
        packed, r, _, _ = self._evaluate()
        order = self._reference_order(packed, None, range(self.num_tracks()), False)
        return r[order].reshape(-1)

    def complete_cost(self):
        _, r, _, _ = self._evaluate()
        return float(np.sum(np.square(r)))

    def residuals_partial(self, camera_ids, track_ids):
        packed, r, _, _ = self._evaluate(camera_ids, track_ids)
        order = self._reference_order(packed, list(camera_ids), list(track_ids), True)
        return r[order].reshape(-1)

    def Jresiduals_partial(self, camera_ids=None, track_ids=None):
        camera_ids, track_ids = list(camera_ids), list(track_ids)
        packed, _, Jc, Jp = self._evaluate(camera_ids, track_ids)
        order = self._reference_order(packed, camera_ids, track_ids, True)
        nc, nt = len(camera_ids), len(track_ids)
        J = np.zeros((2 * len(order), 6 * nc + 3 * nt))
        for row, o in enumerate(order):
            ci, tj = int(packed.obs_cam[o]), int(packed.obs_track[o])
            J[2 * row:2 * row + 2, 6 * ci:6 * ci + 6] = Jc[o]
            J[2 * row:2 * row + 2, 6 * nc + 3 * tj:6 * nc + 3 * tj + 3] = Jp[o]
        return J

    def Jresiduals(self):
        return self.Jresiduals_extended()[0]

    def Jresiduals_extended(self):
        ncam, ntrk = len(self.cameras), self.num_tracks()
        packed, _, Jc, Jp = self._evaluate()
        order = self._reference_order(packed, None, range(ntrk), False)
        J = np.zeros((2 * len(order), self.num_params()))
        row_labels = np.empty((2 * len(order), 2), int)
        col_labels = np.empty((self.num_params(), 2), int)
        for i in range(ncam):
            col_labels[6 * i:6 * i + 6] = (i, -1)
        for j in range(ntrk):
            col_labels[6 * ncam + 3 * j:6 * ncam + 3 * j + 3] = (-1, j)
        for row, o in enumerate(order):
            i, j = int(packed.obs_cam[o]), int(packed.obs_track[o])
            J[2 * row:2 * row + 2, 6 * i:6 * i + 6] = Jc[o]
            J[2 * row:2 * row + 2, 6 * ncam + 3 * j:6 * ncam + 3 * j + 3] = Jp[o]
            row_labels[2 * row:2 * row + 2] = (i, j)
        return J, row_labels, col_labels

    # ---- parameter copies / updates ------------------------------------------------------------
    def clone_params(self):
        """Deep-copy K, cameras and points; share tracks and sensor model (bundle.py:301-310)."""
        b = Bundle()
        b.K = np.array(self.K, dtype=np.float64)
        b.cameras = [Camera(c.R.copy(), c.t.copy(), c.idx) for c in self.cameras]
        b.reconstruction = np.array(self.reconstruction, dtype=np.float64)
        b._tracks = self._tracks
        b._obs_arrays = self._obs_arrays
        b._num_tracks = self._num_tracks
        b.sensor_model = self.sensor_model
        return b

    def perturb(self, delta, param_mask=None):
        delta = np.asarray(delta, dtype=np.float64)
        n = self.num_params()
        if param_mask is not None:
            param_mask = np.asarray(param_mask)
            assert param_mask.shape == (n,), 'shape was ' + str(param_mask.shape)
            full = np.zeros(n)
            full[param_mask] = delta
            delta = full
        assert delta.shape == (n,), 'shape was ' + str(delta.shape)
        for i, cam in enumerate(self.cameras):
            cam.perturb(delta[6 * i:6 * i + 6])
        self.reconstruction = self.reconstruction + delta[6 * len(self.cameras):].reshape(-1, 3)
        return self

    def transform(self, R, t):
        """x -> R x + t on the points, inverse on the cameras (bundle.py:383-395)."""
        R, t = np.asarray(R, float), np.asarray(t, float)
        assert R.shape == (3, 3) and t.shape == (3,)
        self.reconstruction = self.reconstruction.dot(R.T) + t
        for camera in self.cameras:
            camera.transform(R.T, -R.T.dot(t))

    def make_relative_to_first_camera(self):
        R, t = self.cameras[0].Rt
        self.transform(R, t)

    def triangulate(self, track):
        from . import triangulate as _tri
        ids = list(track.camera_ids())
        return _tri.algebraic_lsq(self.K, [self.cameras[i].R for i in ids], [self.cameras[i].t for i in ids],
                                  [track.measurements[i] for i in ids])

    def triangulate_all(self, device=None):
        """Replace every point by its linear triangulation (bundle.py:313-321): all tracks by one
        kernel launch (ba_triangulate) on `device` (default: the current CUDA device).  There is
        no host loop on this path; `Bundle.triangulate(track)` remains as the reference's
        single-track helper."""
        from . import scene as _scene
        cams = list(range(len(self.cameras)))
        trks = list(range(self.num_tracks()))
        if np.shape(self.reconstruction) != (len(trks), 3):
            self.reconstruction = np.zeros((len(trks), 3))
        packed = _scene.pack_scene_device(self, cams, trks, range(len(cams)), range(len(trks)), device)
        prob = _scene.DeviceProblem(packed, device)
        try:
            prob.triangulate()
            _, _, x = prob.download("state")
        finally:
            prob.close()
        self.reconstruction = np.array(x, dtype=np.float64)
