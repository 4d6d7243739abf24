"""CPU oracle (TEST INFRASTRUCTURE — never imported by the product path) for dense bundle
adjustment: a restatement of ba_cuda and the kernels / host code it drives.

  linearise       projective_transform_kernel    src/lib/droid_kernels.cu:176-424
  frame set       torch::_unique(cat(ts, ii))    :1336-1344
  pose blocks A   SparseBlock::update_lhs/rhs     :1131-1173,:1376-1383
  depth terms     C, w, Q with the 0.05 prior     :1396-1400  (accum_cuda :948-998)
  Schur S, v      schur_block / EEt6x6 / Ev6x1    :1222-1311,:1001-1093
  solve           diag += ep + lm*diag; LLT f64; failure => dx = 0   :1192-1213
  back-subst.     EvT6x1 (skips pose index <= 0)  :1095-1115,:1417
  retraction      pose_retr / disp_retr           :898-946,:877-895
All per-pixel math is numpy float32 in the reference's operation order (edge-major, exactly
like the CUDA kernel); the reduced system is assembled and solved in float64 like the
reference's Eigen path.  `dtype=np.float64` switches the per-pixel math to float64 to give an
exact-arithmetic yardstick for tolerance decisions.
Parity pin: cross-checked against the reference's own dense pure-torch BA (src/geom/ba.py,
imported from /root/reference with the oracle SE3 standing in for lietorch) in
tests/test_oracle_pins.py, and on the GPU box against the reference's CUDA kernels (oracle/_ref).
"""
import numpy as np

from . import geom_oracle as G

MIN_DEPTH = 0.25


def linearize(poses, disps, intr, targets, weights, ii, jj, dtype=np.float32):
    """Returns dict with Hs [4,N,6,6], vs [2,N,6], Eii, Eij [N,6,hw], Cii, bz [N,hw]."""
    T = dtype
    poses = np.asarray(poses, np.float32)
    N = len(ii)
    num, ht, wd = disps.shape
    hw = ht * wd
    fx, fy, cx, cy = [T(x) for x in np.asarray(intr, np.float32)]
    t, q = G.edge_pose(poses, ii, jj, stereo_special=True)
    t, q = t.astype(T), q.astype(T)
    stereo = (np.asarray(ii) == np.asarray(jj))[:, None]
    u, v = G.pixel_grid(ht, wd)
    u, v = u.astype(T), v.astype(T)
    di = np.asarray(disps, np.float32)[ii].reshape(N, hw).astype(T)
    Xi = np.stack([np.broadcast_to((u - cx) / fx, (N, hw)), np.broadcast_to((v - cy) / fy, (N, hw)),
                   np.ones((N, hw), T), di], axis=-1)
    if T == np.float32:
        Xj = G.act_se3(t[:, None, :], q[:, None, :], Xi)
    else:
        Xj = _act_se3_any(t[:, None, :], q[:, None, :], Xi)
    x, y, h = Xj[..., 0], Xj[..., 1], Xj[..., 3]
    behind = Xj[..., 2] < T(MIN_DEPTH)
    with np.errstate(divide="ignore"):
        d = np.where(behind, 0.0, 1.0 / Xj[..., 2].astype(np.float64)).astype(T)
    d2 = d * d
    tg = np.asarray(targets, np.float32).reshape(N, 2, hw).astype(T)
    wg = np.asarray(weights, np.float32).reshape(N, 2, hw)
    wu = np.where(behind, 0.0, 0.001 * wg[:, 0].astype(np.float64)).astype(T)
    wv = np.where(behind, 0.0, 0.001 * wg[:, 1].astype(np.float64)).astype(T)
    ru = tg[:, 0] - (fx * d * x + cx)
    rv = tg[:, 1] - (fy * d * y + cy)
    z = np.zeros_like(x)
    Ju = np.stack([fx * (h * d), fx * z, fx * (-x * h * d2), fx * (-x * y * d2),
                   fx * (1 + x * x * d2), fx * (-y * d)], axis=-1).astype(T)
    Jv = np.stack([fy * z, fy * (h * d), fy * (-y * h * d2), fy * (-1 - y * y * d2),
                   fy * (x * y * d2), fy * (x * d)], axis=-1).astype(T)
    Jzu = (fx * (t[:, None, 0] * d - t[:, None, 2] * (x * d2))).astype(T)
    Jzv = (fy * (t[:, None, 1] * d - t[:, None, 2] * (y * d2))).astype(T)
    Cii = wu * Jzu * Jzu + wv * Jzv * Jzv
    bz = wu * ru * Jzu + wv * rv * Jzv
    wu = np.where(stereo, T(0), wu)
    wv = np.where(stereo, T(0), wv)
    adj = G.adj_se3 if T == np.float32 else _adj_se3_any
    Jiu = -adj(t[:, None, :], q[:, None, :], Ju)
    Jiv = -adj(t[:, None, :], q[:, None, :], Jv)
    Jxu = np.concatenate([Jiu, Ju], axis=-1)     # [N,hw,12]
    Jxv = np.concatenate([Jiv, Jv], axis=-1)
    # 12x12 Hessian and gradients, summed over pixels (order differs from the block tree;
    # tests compare with a tolerance that covers fp32 summation order)
    H12 = np.einsum("np,npa,npb->nab", wu, Jxu, Jxu, optimize=True) + \
        np.einsum("np,npa,npb->nab", wv, Jxv, Jxv, optimize=True)
    g12 = np.einsum("np,npa->na", wu * ru, Jxu, optimize=True) + \
        np.einsum("np,npa->na", wv * rv, Jxv, optimize=True)
    Hs = np.stack([H12[:, :6, :6], H12[:, :6, 6:], H12[:, 6:, :6], H12[:, 6:, 6:]], axis=0)
    vs = np.stack([g12[:, :6], g12[:, 6:]], axis=0)
    Eii = ((wu * Jzu)[..., None] * Jiu + (wv * Jzv)[..., None] * Jiv).transpose(0, 2, 1)
    Eij = ((wu * Jzu)[..., None] * Ju + (wv * Jzv)[..., None] * Jv).transpose(0, 2, 1)
    return dict(Hs=Hs, vs=vs, Eii=Eii, Eij=Eij, Cii=Cii, bz=bz)


def _act_so3_any(q, X):
    u = 2.0 * np.cross(np.broadcast_to(q[..., :3], X.shape), X)
    return X + q[..., 3:4] * u + np.cross(np.broadcast_to(q[..., :3], X.shape), u)


def _act_se3_any(t, q, X):
    Y = _act_so3_any(q, X[..., :3]) + X[..., 3:4] * t
    return np.concatenate([Y, X[..., 3:4]], axis=-1)


def _adj_se3_any(t, q, X):
    qi = np.concatenate([-q[..., :3], q[..., 3:]], axis=-1)
    a, b = X[..., :3], X[..., 3:]
    u = np.cross(a, np.broadcast_to(t, a.shape))
    return np.concatenate([_act_so3_any(qi, a), _act_so3_any(qi, b) + _act_so3_any(qi, u)], axis=-1)


def phase1(poses, disps, intrinsics, disps_sens, targets, weights, eta, ii, jj, t0, t1, motion_only,
           dtype=np.float32):
    """Linearise the given (local) edges and build their contribution to the reduced camera
    system (A - S | b - v) in float64, plus what the back-substitution needs."""
    T = dtype
    ii = np.asarray(ii, np.int64)
    jj = np.asarray(jj, np.int64)
    N = len(ii)
    num, ht, wd = np.asarray(disps).shape
    hw = ht * wd
    P = t1 - t0
    n = 6 * P
    ts = np.arange(t0, t1)
    ii_exp = np.concatenate([ts, ii])
    jj_exp = np.concatenate([ts, jj])
    kx, kk_exp = np.unique(ii_exp, return_inverse=True)
    M = len(kx)
    lin = linearize(poses, disps, intrinsics, targets, weights, ii, jj, dtype=T) if N > 0 else None
    # ---- pose x pose block A (float64, like the Eigen triplets)
    A = np.zeros((n, n), np.float64)
    b = np.zeros(n, np.float64)
    for kblk, (ri, ci) in enumerate([(ii, ii), (ii, jj), (jj, ii), (jj, jj)]):
        for e in range(N):
            r, c = ri[e] - t0, ci[e] - t0
            if r >= 0 and c >= 0 and r < P and c < P:
                A[6 * r:6 * r + 6, 6 * c:6 * c + 6] += lin["Hs"][kblk, e].astype(np.float32).astype(np.float64)
    for kblk, ri in enumerate((ii, jj)):
        for e in range(N):
            r = ri[e] - t0
            if 0 <= r < P:
                b[6 * r:6 * r + 6] += lin["vs"][kblk, e].astype(np.float32).astype(np.float64)
    st = dict(A=A, kx=kx, kk_exp=kk_exp, jj_exp=jj_exp, P=P, N=N, lin=lin)
    if motion_only:
        st.update(Hred=A, bred=b)
        return st
    disps = np.asarray(disps, np.float32)
    disps_sens = np.asarray(disps_sens, np.float32)
    alpha = T(0.05)
    m = (disps_sens[kx].reshape(M, hw) > 0).astype(T)
    eta_a = np.asarray(eta, np.float32).reshape(-1, hw).astype(T)
    if eta_a.shape[0] == num:          # frame-indexed eta (sharded callers): pick the rows of kx
        eta_a = eta_a[kx]
    eta_m = np.broadcast_to(eta_a, (M, hw))
    C = np.zeros((M, hw), T)
    w = np.zeros((M, hw), T)
    for e in range(N):
        k = kk_exp[P + e]
        C[k] += lin["Cii"][e]
        w[k] += lin["bz"][e]
    C = C + m * alpha + (1 - m) * eta_m
    w = w - m * alpha * (disps[kx].reshape(M, hw).astype(T) - disps_sens[kx].reshape(M, hw).astype(T))
    Q = (T(1.0) / C).astype(T)
    Ei = np.zeros((P, 6, hw), T)
    for e in range(N):
        p = ii[e] - t0
        if 0 <= p < P:
            Ei[p] += lin["Eii"][e]
    E = np.concatenate([Ei, lin["Eij"].astype(T)], axis=0) if N > 0 else Ei     # [P+N,6,hw]
    # ---- Schur complement (schur_block): entries n with t0 <= jj_exp[n] < t1
    S = np.zeros((n, n), np.float64)
    v = np.zeros(n, np.float64)
    for k in range(M):
        ent = [a for a in range(P + N) if kk_exp[a] == k and t0 <= jj_exp[a] < t1]
        for a in ent:
            pa = jj_exp[a] - t0
            EaQ = E[a] * Q[k][None]
            v[6 * pa:6 * pa + 6] += (EaQ * w[k][None]).sum(-1).astype(np.float32).astype(np.float64)
            for bb in ent:
                pb = jj_exp[bb] - t0
                S[6 * pa:6 * pa + 6, 6 * pb:6 * pb + 6] += (EaQ @ E[bb].T).astype(np.float32).astype(np.float64)
    st.update(Hred=A - S, bred=b - v, E=E, Q=Q, w=w, C=C)
    return st


def solve(Hred, bred, A, lm, ep, damping="reduced"):
    """diag += ep + lm*diag, float64 LLT, failure => zero step (:1192-1213)."""
    n = Hred.shape[0]
    L = Hred.copy()
    dg = np.diag(L).copy() if damping == "reduced" else np.diag(A).copy()
    L[np.diag_indices(n)] = np.diag(L) + np.float64(np.float32(ep)) + np.float64(np.float32(lm)) * dg
    try:
        c = np.linalg.cholesky(L)
        x = np.linalg.solve(c.T, np.linalg.solve(c, bred))
        ok = bool(np.all(np.isfinite(x)))
    except np.linalg.LinAlgError:
        ok = False
    P = n // 6
    return (x.reshape(P, 6).astype(np.float32) if ok else np.zeros((P, 6), np.float32)), (0 if ok else 1)


def phase2(st, dx, poses, disps, t0, t1, motion_only, owner_lo=0, owner_hi=None):
    """Back-substitute dz for the frames in [owner_lo, owner_hi) and retract.  In place on copies
    the caller owns; returns dz indexed by frame."""
    num, ht, wd = disps.shape
    hw = ht * wd
    owner_hi = num if owner_hi is None else owner_hi
    P, N, kx, kk_exp, jj_exp = st["P"], st["N"], st["kx"], st["kk_exp"], st["jj_exp"]
    dz_frames = np.zeros((num, hw), np.float32)
    if not motion_only:
        E, Q, w = st["E"], st["Q"], st["w"]
        T = E.dtype.type
        ix = jj_exp - t0
        acc = np.zeros_like(Q)
        for a in range(P + N):
            if 0 < ix[a] < P:                     # EvT6x1 skips pose index <= 0 (:1105)
                acc[kk_exp[a]] += (E[a] * dx[ix[a]].astype(T)[:, None]).sum(0)
        dz = (Q * (w - acc)).astype(np.float32)
        own = (kx >= owner_lo) & (kx < owner_hi)
        d = disps.reshape(num, hw)
        d[kx[own]] = d[kx[own]] + dz[own]
        dz_frames[kx[own]] = dz[own]
    tn, qn = G.retr_se3(dx, poses[t0:t1, :3], poses[t0:t1, 3:])
    poses[t0:t1, :3] = tn
    poses[t0:t1, 3:] = qn
    return dz_frames


def ba(poses, disps, intrinsics, disps_sens, targets, weights, eta, ii, jj, t0, t1, iterations,
       lm, ep, motion_only, dtype=np.float32, return_debug=False, damping="reduced"):
    """In-place-free restatement of droid_backends.ba: returns (poses, disps, dx, dz_by_frame, status).
    damping="reduced" is the CUDA path (diag(A-S) += ep + lm*diag(A-S), :1196-1197);
    damping="pose_block" is the convention of the reference's pure-torch BA (src/geom/chol.py:56-57:
    damp H before subtracting E Q E^T) and exists only to cross-check against that code."""
    poses = np.array(poses, np.float32, copy=True)
    disps = np.array(disps, np.float32, copy=True)
    num, ht, wd = disps.shape
    P = t1 - t0
    dx = np.zeros((P, 6), np.float32)
    dz_frames = np.zeros((num, ht * wd), np.float32)
    status, debug = [], {}
    for _ in range(iterations):
        st = phase1(poses, disps, intrinsics, disps_sens, targets, weights, eta, ii, jj, t0, t1, motion_only, dtype)
        dx, fail = solve(st["Hred"], st["bred"], st["A"], lm, ep, damping)
        status.append(fail)
        if return_debug:
            debug = {k: st[k] for k in ("Hred", "bred", "lin", "E", "Q", "w", "C") if k in st}
        dz_frames = phase2(st, dx, poses, disps, t0, t1, motion_only)
    out = (poses, disps, dx, dz_frames, np.asarray(status, np.int32))
    return out + (debug,) if return_debug else out


def reprojection_cost(poses, disps, intr, targets, weights, ii, jj):
    """0.5 * sum w r^2 over all edges (property test: BA must not increase it much)."""
    lin_in = dict(poses=poses, disps=disps)
    N = len(ii)
    num, ht, wd = np.asarray(disps).shape
    hw = ht * wd
    fx, fy, cx, cy = [np.float64(x) for x in np.asarray(intr)]
    t, q = G.edge_pose(np.asarray(poses, np.float32), ii, jj, stereo_special=True)
    u, v = G.pixel_grid(ht, wd)
    di = np.asarray(disps, np.float64)[ii].reshape(N, hw)
    Xi = np.stack([np.broadcast_to((u - cx) / fx, (N, hw)), np.broadcast_to((v - cy) / fy, (N, hw)),
                   np.ones((N, hw)), di], axis=-1)
    Xj = _act_se3_any(t[:, None, :].astype(np.float64), q[:, None, :].astype(np.float64), Xi)
    ok = Xj[..., 2] >= MIN_DEPTH
    with np.errstate(divide="ignore", invalid="ignore"):
        pu = fx * Xj[..., 0] / Xj[..., 2] + cx
        pv = fy * Xj[..., 1] / Xj[..., 2] + cy
    tg = np.asarray(targets, np.float64).reshape(N, 2, hw)
    wg = np.asarray(weights, np.float64).reshape(N, 2, hw)
    r2 = np.where(ok, wg[:, 0] * (tg[:, 0] - pu) ** 2 + wg[:, 1] * (tg[:, 1] - pv) ** 2, 0.0)
    del lin_in
    return 0.5 * float(r2.sum())
