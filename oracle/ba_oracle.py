"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy/scipy, FP64) of pysfm's
Schur-complement Levenberg-Marquardt bundle adjuster.

This file is the *checker* for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``pysfm_b200/`` does, and the product never falls back to it.

It restates the reference's algorithm in observation-list form (one row per measurement,
point-major), which is what makes BASELINE configs 2/4/5 feasible on a CPU at all: the
reference's own loops visit all (track, camera, camera) triples and allocate a dense
(nc, nt, 6, 3) array (bundle_adjuster.py:107, :267-276).  Every function cites the reference
lines it follows.

PARITY PINNED: ``tests/test_oracle_golden.py`` checks every stage of this file against
fixtures in ``tests/golden/`` that were produced by the UNMODIFIED reference, loaded through
``oracle/refshim.py`` by ``oracle/make_golden.py`` (committed).  When ``/root/reference`` is
present the same test also re-runs the reference live.  Agreement: <= 1e-12 relative on every
intermediate (HCCs, HPPs, HCPs, bCs, bPs, HPP_invs, S, b, motion/structure update, cost) and on
the accepted-cost sequence of ``optimize``.
"""
import numpy as np

try:
    import scipy.sparse as _sp
except Exception:  # pragma: no cover - scipy is part of the image
    _sp = None


class Problem(object):
    """Plain arrays describing the sub-problem ``BundleAdjuster.set_bundle`` selected
    (bundle_adjuster.py:54-111).

    K (3,3); R (nc,3,3); t (nc,3); x (nt,3)              cameras / points by *position*
    obs_cam (nobs,), obs_pt (nobs,), obs_uv (nobs,2)      measurements of selected tracks in
                                                          selected cameras, any order
    model: ('gaussian', L (2,2)) or ('cauchy', sigma)
    optim_cam: positions of optimised cameras, in reduced-system order (optim_camera_indices)
    optim_pt:  positions of updated tracks (optim_track_indices)
    """

    def __init__(self, K, R, t, x, obs_cam, obs_pt, obs_uv, model, optim_cam, optim_pt):
        self.K = np.asarray(K, dtype=np.float64).reshape(3, 3)
        self.R = np.asarray(R, dtype=np.float64).reshape(-1, 3, 3)
        self.t = np.asarray(t, dtype=np.float64).reshape(-1, 3)
        self.x = np.asarray(x, dtype=np.float64).reshape(-1, 3)
        self.obs_cam = np.asarray(obs_cam, dtype=np.int64)
        self.obs_pt = np.asarray(obs_pt, dtype=np.int64)
        self.obs_uv = np.asarray(obs_uv, dtype=np.float64).reshape(-1, 2)
        self.model = model
        self.optim_cam = np.asarray(optim_cam, dtype=np.int64)
        self.optim_pt = np.asarray(optim_pt, dtype=np.int64)

    @property
    def nc(self):
        return len(self.R)

    @property
    def nt(self):
        return len(self.x)

    def with_params(self, R, t, x):
        return Problem(self.K, R, t, x, self.obs_cam, self.obs_pt, self.obs_uv, self.model,
                       self.optim_cam, self.optim_pt)

    def cam_slot(self):
        s = np.full(self.nc, -1, dtype=np.int64)
        s[self.optim_cam] = np.arange(len(self.optim_cam))
        return s

    def pt_slot(self):
        s = np.full(self.nt, -1, dtype=np.int64)
        s[self.optim_pt] = np.arange(len(self.optim_pt))
        return s


# --------------------------------------------------------------------------------------------
def gaussian_L(cov):
    """GaussianModel.__init__ (sensor_model.py:8-17): L = chol(inv(cov)), lower."""
    c = np.asarray(cov, dtype=np.float64)
    if c.ndim == 0:
        c = float(c) * np.eye(2)
    elif c.shape == (2,):
        c = np.diag(c)
    return np.linalg.cholesky(np.linalg.inv(c))


def so3_exp(m):
    """lie.py:21-34 -- Rodrigues, identity below 1e-8."""
    m = np.asarray(m, dtype=np.float64)
    th = np.linalg.norm(m)
    if th < 1e-8:
        return np.eye(3)
    W = np.array([[0., -m[2], m[1]], [m[2], 0., -m[0]], [-m[1], m[0], 0.]])
    return np.eye(3) + (np.sin(th) / th) * W + ((1. - np.cos(th)) / (th * th)) * W.dot(W)


def sensor(model, e):
    """residual_from_error / Jresidual_from_error for all observations at once.
    Gaussian: sensor_model.py:23-29.  Cauchy: sensor_model.py:48-69."""
    kind = model[0]
    n = len(e)
    if kind == 'gaussian':
        L = np.asarray(model[1], dtype=np.float64)
        return e.dot(L.T), np.broadcast_to(L, (n, 2, 2))
    sigma = float(model[1])
    rho = np.sqrt(np.sum(e * e, axis=1))
    small = rho < 1e-5
    safe = np.where(small, 1.0, rho)
    s = np.sqrt(np.log(1. + safe * safe / (sigma * sigma)))
    r = e * (s / safe)[:, None]
    ee = e[:, :, None] * e[:, None, :]
    eye = np.eye(2)[None]
    with np.errstate(divide='ignore', invalid='ignore'):
        J = ee / (safe * s * (safe * safe + sigma * sigma))[:, None, None] + \
            (safe[:, None, None] * eye - ee / safe[:, None, None]) * (s / (safe * safe))[:, None, None]
    r[small] = e[small] / sigma
    J[small] = np.eye(2) / sigma
    return r, J


def linearize(P):
    """Per observation: residual r (n,2), Jc (n,2,6), Jp (n,2,3).
    bundle.py:243-252 (residual) and :255-277 (Jresidual); Jpr bundle.py:8-11;
    J_expm_x = skew(-x) lie.py:38-40; camera parameter order [rotation | translation]."""
    R = P.R[P.obs_cam]
    x = P.x[P.obs_pt]
    y = np.einsum('nij,nj->ni', R, x) + P.t[P.obs_cam]
    p = y.dot(P.K.T)
    pred = p[:, :2] / p[:, 2:3]
    e = pred - P.obs_uv
    n = len(e)
    Jpr = np.zeros((n, 2, 3))
    Jpr[:, 0, 0] = 1. / p[:, 2]
    Jpr[:, 1, 1] = 1. / p[:, 2]
    Jpr[:, 0, 2] = -p[:, 0] / (p[:, 2] * p[:, 2])
    Jpr[:, 1, 2] = -p[:, 1] / (p[:, 2] * p[:, 2])
    Jt = Jpr.dot(P.K)                              # (n,2,3)
    Jx = np.einsum('nij,njk->nik', Jt, R)
    skew_neg = np.zeros((n, 3, 3))                 # skew(-x)
    skew_neg[:, 0, 1] = x[:, 2]
    skew_neg[:, 0, 2] = -x[:, 1]
    skew_neg[:, 1, 0] = -x[:, 2]
    skew_neg[:, 1, 2] = x[:, 0]
    skew_neg[:, 2, 0] = x[:, 1]
    skew_neg[:, 2, 1] = -x[:, 0]
    JR = np.einsum('nij,njk->nik', Jx, skew_neg)
    r, Jr = sensor(P.model, e)
    Jc = np.einsum('nij,njk->nik', Jr, np.concatenate((JR, Jt), axis=2))
    Jp = np.einsum('nij,njk->nik', Jr, Jx)
    return r, Jc, Jp


def prepare(P):
    """prepare_schur_complement (bundle_adjuster.py:211-234): HCCs (nc,6,6), HPPs (nt,3,3),
    per-observation HCP blocks W (nobs,6,3), bCs (nc,6), bPs (nt,3)."""
    r, Jc, Jp = linearize(P)
    HCCs = np.zeros((P.nc, 6, 6))
    HPPs = np.zeros((P.nt, 3, 3))
    bCs = np.zeros((P.nc, 6))
    bPs = np.zeros((P.nt, 3))
    np.add.at(HCCs, P.obs_cam, np.einsum('nai,naj->nij', Jc, Jc))
    np.add.at(HPPs, P.obs_pt, np.einsum('nai,naj->nij', Jp, Jp))
    np.add.at(bCs, P.obs_cam, np.einsum('nai,na->ni', Jc, r))
    np.add.at(bPs, P.obs_pt, np.einsum('nai,na->ni', Jp, r))
    W = np.einsum('nai,naj->nij', Jc, Jp)
    return dict(HCCs=HCCs, HPPs=HPPs, W=W, bCs=bCs, bPs=bPs, r=r, Jc=Jc, Jp=Jp)


def dense_HCPs(P, W):
    """The reference's dense (nc, nt, 6, 3) array (bundle_adjuster.py:107,232)."""
    out = np.zeros((P.nc, P.nt, 6, 3))
    out[P.obs_cam, P.obs_pt] = W
    return out


def apply_damping(blocks, damping):
    """apply_damping (bundle_adjuster.py:238-242) -> optimize.apply_lm_damping_inplace
    (optimize.py:7-9): diagonal *= (1 + damping) on every camera and point block."""
    d6, d3 = np.arange(6), np.arange(3)
    blocks['HCCs'][:, d6, d6] *= (1. + damping)
    blocks['HPPs'][:, d3, d3] *= (1. + damping)


def schur(P, blocks, rcond=1e-5):
    """compute_schur_complement (bundle_adjuster.py:247-278).  Returns S (nc',nc',6,6),
    b (nc',6) and HPP_invs (nt,3,3); pinv with relative cutoff ``rcond`` (:256) or inv (:254)."""
    if rcond is None:
        Vinv = np.linalg.inv(blocks['HPPs'])
    else:
        Vinv = np.linalg.pinv(blocks['HPPs'], rcond)
    slot = P.cam_slot()
    nco = len(P.optim_cam)
    n = 6 * nco
    keep = slot[P.obs_cam] >= 0
    W = blocks['W'][keep]
    oc = slot[P.obs_cam[keep]]
    op = P.obs_pt[keep]
    A = np.zeros((n, n))
    rhs = np.zeros(n)
    for pos, i in enumerate(P.optim_cam):                         # :263-265
        A[6 * pos:6 * pos + 6, 6 * pos:6 * pos + 6] = blocks['HCCs'][i]
        rhs[6 * pos:6 * pos + 6] = blocks['bCs'][i]
    if len(W):
        Y = np.einsum('nij,njk->nik', W, Vinv[op])                # W_ik V_k^-1
        rows = (6 * oc[:, None, None] + np.arange(6)[None, :, None]) + np.zeros((1, 1, 3), dtype=np.int64)
        cols = (3 * op[:, None, None] + np.arange(3)[None, None, :]) + np.zeros((1, 6, 1), dtype=np.int64)
        if _sp is not None:
            Ym = _sp.csr_matrix((Y.ravel(), (rows.ravel(), cols.ravel())), shape=(n, 3 * P.nt))
            Wm = _sp.csr_matrix((W.ravel(), (rows.ravel(), cols.ravel())), shape=(n, 3 * P.nt))
            A -= (Ym @ Wm.T).toarray()                            # :276
            rhs -= Ym @ blocks['bPs'].reshape(-1)                 # :274
        else:
            Yd = np.zeros((n, 3 * P.nt))
            Wd = np.zeros((n, 3 * P.nt))
            Yd[rows.ravel(), cols.ravel()] = Y.ravel()
            Wd[rows.ravel(), cols.ravel()] = W.ravel()
            A -= Yd.dot(Wd.T)
            rhs -= Yd.dot(blocks['bPs'].reshape(-1))
    S = A.reshape(nco, 6, nco, 6).transpose(0, 2, 1, 3).copy()
    return S, rhs.reshape(nco, 6), Vinv


def solve_motion(S, b, cam_param_mask=None):
    """solve_motion_normal_eqns (bundle_adjuster.py:281-312): flatten, mask, LU solve,
    scatter back with zeros.  numpy.linalg.LinAlgError propagates (-> ill-conditioned)."""
    nco = S.shape[0]
    A = S.transpose(0, 2, 1, 3).reshape(6 * nco, 6 * nco)
    rhs = b.reshape(-1)
    if cam_param_mask is None:
        cam_param_mask = np.ones(6 * nco, bool)
    m = np.asarray(cam_param_mask, dtype=bool)
    sol = np.linalg.solve(A[m][:, m], rhs[m])
    dC = np.zeros(6 * nco)
    dC[m] = sol
    return dC.reshape(nco, 6)


def backsubstitute(P, blocks, Vinv, dC):
    """backsubstitute (bundle_adjuster.py:316-331) for the optimised tracks."""
    slot = P.cam_slot()
    keep = slot[P.obs_cam] >= 0
    acc = np.zeros((P.nt, 3))
    if keep.any():
        contrib = np.einsum('nij,ni->nj', blocks['W'][keep], dC[slot[P.obs_cam[keep]]])
        np.add.at(acc, P.obs_pt[keep], contrib)
    dP = np.einsum('nij,nj->ni', Vinv, blocks['bPs'] - acc)
    return dP[P.optim_pt]


def compute_update(P, damping, cam_param_mask=None, rcond=1e-5):
    """compute_update (bundle_adjuster.py:176-208): returns (-dC, -dP)."""
    blocks = prepare(P)
    apply_damping(blocks, damping)
    S, b, Vinv = schur(P, blocks, rcond)
    dC = solve_motion(S, b, cam_param_mask)
    dP = backsubstitute(P, blocks, Vinv, dC)
    return -dC, -dP


def compute_cost(P):
    """compute_cost (bundle_adjuster.py:165-171): optimised tracks x optimised cameras only."""
    r, _, _ = linearize(P)
    keep = (P.cam_slot()[P.obs_cam] >= 0) & (P.pt_slot()[P.obs_pt] >= 0)
    return float(np.sum(r[keep] ** 2))


def apply_update(P, motion, structure):
    """update_motion / update_structure (bundle_adjuster.py:334-343) on a copy:
    R <- R exp(d[:3]), t <- t + d[3:] (bundle.py:76-80); x <- x + d."""
    R, t, x = P.R.copy(), P.t.copy(), P.x.copy()
    for pos, i in enumerate(P.optim_cam):
        R[i] = R[i].dot(so3_exp(motion[pos, :3]))
        t[i] = t[i] + motion[pos, 3:]
    x[P.optim_pt] += structure
    return P.with_params(R, t, x)


def optimize(P, cam_param_mask=None, max_steps=25, init_damping=10., improvement_threshold=1e-4,
             rcond=1e-5):
    """optimize (bundle_adjuster.py:117-162).  Returns (final problem, info dict)."""
    damping = init_damping
    num_steps, converged = 0, False
    costs = [compute_cost(P)]
    trace = []
    while not converged and num_steps < max_steps:
        num_steps += 1
        cur_cost = compute_cost(P)
        while not converged and damping < 1e+8:
            try:
                motion, structure = compute_update(P, damping, cam_param_mask, rcond)
            except np.linalg.LinAlgError:
                damping *= 10.
                converged = damping > 1e+8
                continue
            Pn = apply_update(P, motion, structure)
            next_cost = compute_cost(Pn)
            trace.append(dict(step=num_steps, damping=damping, cost=cur_cost, cand_cost=next_cost,
                              accepted=bool(next_cost < cur_cost)))
            if next_cost < cur_cost:
                damping *= .1
                P = Pn
                costs.append(next_cost)
                converged = abs(cur_cost - next_cost) < improvement_threshold
                break
            else:
                damping *= 10.
                converged = damping > 1e+8
    return P, dict(costs=costs, num_steps=num_steps, converged=converged, trace=trace)


def lm_iteration(P, damping, rcond=1e-5):
    """The unit bench.py times: one compute_update + one candidate compute_cost."""
    motion, structure = compute_update(P, damping, None, rcond)
    return compute_cost(apply_update(P, motion, structure))


def reprojection_rmse(P):
    """sqrt(mean ||pred - z||^2) in pixels over all selected observations."""
    y = np.einsum('nij,nj->ni', P.R[P.obs_cam], P.x[P.obs_pt]) + P.t[P.obs_cam]
    p = y.dot(P.K.T)
    e = p[:, :2] / p[:, 2:3] - P.obs_uv
    return float(np.sqrt(np.mean(np.sum(e * e, axis=1))))
