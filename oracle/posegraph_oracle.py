"""ORACLE — test infrastructure only (see oracle/icp_oracle.cpp header).  Pose-graph side.

CPU restatement (numpy + scipy.sparse) of what laser_slam asks of GTSAM on the hot path:
  * factors built by LaserTrack::makeMeasurementFactor / makeRelativeMeasurementFactor
    (reference laser_slam/src/laser_track.cpp:431-458): prior  h(x) = T_w,  between  h(x) = T_w_a^-1 * T_w_b
    (optionally with node a frozen, laser_track.cpp:440-444), error = Local(measured, h(x)) in R^6;
  * noise models Diagonal::Sigmas(6) and Robust(Cauchy(1), Diagonal) (laser_track.cpp:37-64,
    incremental_estimator.cpp:29-48);
  * IncrementalEstimator::estimate = three Gauss-Newton passes over the graph
    (incremental_estimator.cpp:151-163; iSAM2 with relinearizeSkip 1 / threshold 1e-3 relinearises essentially
    everything every pass, so the batch GN fixed point is what both sides converge to).

PARITY UNPINNED: GTSAM (gtborg/gtsam @ b66dda2f..., dependencies.rosinstall:42-45), minkindr and minkindr_gtsam
are absent and the reference holds no numeric test for this path.  [DEFINED] conventions:
  tangent / residual order  [translation(3); rotation-vector(3)]   (sigmas [0.005 m x3, 0.0015 rad x3])
  Local(M, X) for SE3       [ R_M^T (t_X - t_M) ;  Log(R_M^T R_X) ]  (decoupled chart, as kindr::minimal)
  retraction                t += dt ;  R <- R * Exp(dr)
  Cauchy(k=1)               IRLS weight  w = 1 / (1 + ||r/sigma||^2), Jacobian and residual scaled by sqrt(w)
The Gauss-Newton fixed point depends only on the residual definition, not on the retraction.
Poses are rows [qw, qx, qy, qz, tx, ty, tz] (Hamilton quaternion), the layout of the C ABI (ls_pg_*).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

PRIOR, BETWEEN = 0, 1


# ---------------------------------------------------------------- SO3 / SE3 helpers (batched, float64)
def quat_to_R(q):
    q = np.asarray(q, np.float64)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - w * z); R[..., 0, 2] = 2 * (x * z + w * y)
    R[..., 1, 0] = 2 * (x * y + w * z); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - w * x)
    R[..., 2, 0] = 2 * (x * z - w * y); R[..., 2, 1] = 2 * (y * z + w * x); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def R_to_quat(R):
    R = np.asarray(R, np.float64)
    out = np.empty(R.shape[:-2] + (4,))
    flat_R = R.reshape(-1, 3, 3)
    flat = out.reshape(-1, 4)
    for i, m in enumerate(flat_R):
        t = m[0, 0] + m[1, 1] + m[2, 2]
        if t > 0:
            s = np.sqrt(t + 1.0) * 2
            q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
        elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
            s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
            q = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
        elif m[1, 1] > m[2, 2]:
            s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
            q = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
        else:
            s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
            q = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
        q = np.array(q)
        flat[i] = q / np.linalg.norm(q) * (1.0 if q[0] >= 0 else -1.0)
    return out


def skew(v):
    v = np.asarray(v, np.float64)
    S = np.zeros(v.shape[:-1] + (3, 3))
    S[..., 0, 1] = -v[..., 2]; S[..., 0, 2] = v[..., 1]
    S[..., 1, 0] = v[..., 2]; S[..., 1, 2] = -v[..., 0]
    S[..., 2, 0] = -v[..., 1]; S[..., 2, 1] = v[..., 0]
    return S


def so3_exp(w):
    w = np.asarray(w, np.float64)
    th = np.linalg.norm(w, axis=-1)[..., None, None]
    K = skew(w)
    small = th < 1e-8
    th_s = np.where(small, 1.0, th)
    a = np.where(small, 1.0 - th * th / 6.0, np.sin(th_s) / th_s)
    b = np.where(small, 0.5 - th * th / 24.0, (1.0 - np.cos(th_s)) / (th_s * th_s))
    return np.eye(3) + a * K + b * (K @ K)


def so3_log(R):
    R = np.asarray(R, np.float64)
    v = 0.5 * np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], -1)
    s = np.linalg.norm(v, axis=-1)
    c = 0.5 * (np.trace(R, axis1=-2, axis2=-1) - 1.0)
    th = np.arctan2(s, c)
    small = s < 1e-8
    scale = np.where(small, 1.0 + th * th / 6.0, th / np.where(small, 1.0, s))
    # (rotations near pi do not occur on this path: relative poses between consecutive scans / loop closures)
    return v * scale[..., None]


def jr_inv(phi):
    """Inverse right Jacobian of SO3: d Log(R Exp(d)) / d d at d = 0, R = Exp(phi)."""
    phi = np.asarray(phi, np.float64)
    th = np.linalg.norm(phi, axis=-1)[..., None, None]
    K = skew(phi)
    small = th < 1e-6
    th_s = np.where(small, 1.0, th)
    c = np.where(small, 1.0 / 12.0, 1.0 / (th_s * th_s) - (1.0 + np.cos(th_s)) / (2.0 * th_s * np.sin(th_s)))
    return np.eye(3) + 0.5 * K + c * (K @ K)


def se3_compose(A, B):
    RA, RB = quat_to_R(A[..., :4]), quat_to_R(B[..., :4])
    R = RA @ RB
    t = (RA @ B[..., 4:, None])[..., 0] + A[..., 4:]
    return np.concatenate([R_to_quat(R), t], -1)


def se3_inverse(A):
    R = quat_to_R(A[..., :4])
    Rt = np.swapaxes(R, -1, -2)
    return np.concatenate([R_to_quat(Rt), -(Rt @ A[..., 4:, None])[..., 0]], -1)


def se3_from_matrix(T):
    T = np.asarray(T, np.float64)
    return np.concatenate([R_to_quat(T[..., :3, :3]), T[..., :3, 3]], -1)


def se3_to_matrix(P):
    P = np.asarray(P, np.float64)
    T = np.zeros(P.shape[:-1] + (4, 4))
    T[..., :3, :3] = quat_to_R(P[..., :4])
    T[..., :3, 3] = P[..., 4:]
    T[..., 3, 3] = 1.0
    return T


# ---------------------------------------------------------------- factors
def make_factor(ftype, key_a, key_b, meas7, sigma6, robust=0, fix_a=0, fixed_a7=None):
    return dict(type=int(ftype), key_a=int(key_a), key_b=int(key_b), meas=np.asarray(meas7, np.float64),
                sigma=np.asarray(sigma6, np.float64), robust=int(robust), fix_a=int(fix_a),
                fixed_a=np.asarray(fixed_a7 if fixed_a7 is not None else [1, 0, 0, 0, 0, 0, 0], np.float64))


def linearize(factors, keys, poses):
    """Whitened, robust-weighted residuals and Jacobians.  Returns r (F,6), Ja (F,6,6), Jb (F,6,6), ia, ib
    (pose indices, ia = -1 when the factor does not depend on node a), and the robust cost."""
    index = {int(k): i for i, k in enumerate(keys)}
    F = len(factors)
    r = np.zeros((F, 6)); Ja = np.zeros((F, 6, 6)); Jb = np.zeros((F, 6, 6))
    ia = np.full(F, -1, np.int64); ib = np.zeros(F, np.int64)
    cost = 0.0
    for f, fac in enumerate(factors):
        Rm, tm = quat_to_R(fac["meas"][:4]), fac["meas"][4:]
        if fac["type"] == PRIOR:
            ib[f] = index[fac["key_a"]]
            X = poses[ib[f]]
            R, t = quat_to_R(X[:4]), X[4:]
            rt = Rm.T @ (t - tm)
            rR = so3_log(Rm.T @ R)
            jb = np.zeros((6, 6))
            jb[:3, :3] = Rm.T
            jb[3:, 3:] = jr_inv(rR)
            ja = np.zeros((6, 6))
        else:
            ib[f] = index[fac["key_b"]]
            B = poses[ib[f]]
            if fac["fix_a"]:
                A = fac["fixed_a"]
            else:
                ia[f] = index[fac["key_a"]]
                A = poses[ia[f]]
            Ra, ta, Rb, tb = quat_to_R(A[:4]), A[4:], quat_to_R(B[:4]), B[4:]
            v = Ra.T @ (tb - ta)
            rt = Rm.T @ (v - tm)
            RE = Rm.T @ Ra.T @ Rb
            rR = so3_log(RE)
            Ji = jr_inv(rR)
            ja = np.zeros((6, 6)); jb = np.zeros((6, 6))
            ja[:3, :3] = -Rm.T @ Ra.T
            ja[:3, 3:] = Rm.T @ skew(v)
            ja[3:, 3:] = -Ji @ Rb.T @ Ra
            jb[:3, :3] = Rm.T @ Ra.T
            jb[3:, 3:] = Ji
            if fac["fix_a"]:
                ja[:] = 0.0
        res = np.concatenate([rt, rR]) / fac["sigma"]
        ja = ja / fac["sigma"][:, None]
        jb = jb / fac["sigma"][:, None]
        e2 = float(res @ res)
        if fac["robust"]:
            w = 1.0 / (1.0 + e2)
            cost += 0.5 * np.log1p(e2)          # Cauchy rho with k = 1
            sw = np.sqrt(w)
            res, ja, jb = res * sw, ja * sw, jb * sw
        else:
            cost += 0.5 * e2
        r[f], Ja[f], Jb[f] = res, ja, jb
    return r, Ja, Jb, ia, ib, cost


def retract(poses, delta):
    out = poses.copy()
    out[:, 4:] += delta[:, :3]
    R = quat_to_R(poses[:, :4]) @ so3_exp(delta[:, 3:])
    out[:, :4] = R_to_quat(R)
    return out


def gauss_newton_step(factors, keys, poses, damp_keys=()):
    """One Gauss-Newton step.  damp_keys: poses that get the gauge damping diag(1 x3, 4 x3) added to H only
    ([DEFINED]: first pose of a track whose prior was removed after linking, incremental_estimator.cpp:212-237;
    the fixed point is unchanged because the gradient is untouched)."""
    r, Ja, Jb, ia, ib, cost = linearize(factors, keys, poses)
    P = len(keys)
    rows, cols, vals = [], [], []
    index = {int(k): i for i, k in enumerate(keys)}
    for dk in damp_keys:
        i = index[int(dk)]
        rows.append(6 * i + np.arange(6)); cols.append(6 * i + np.arange(6)); vals.append(np.array([1.0] * 3 + [4.0] * 3))
    g = np.zeros(6 * P)
    blk = np.arange(6)
    for f in range(len(factors)):
        nodes = [(ib[f], Jb[f])] + ([(ia[f], Ja[f])] if ia[f] >= 0 else [])
        for (i, Ji) in nodes:
            g[6 * i:6 * i + 6] += Ji.T @ r[f]
            for (j, Jj) in nodes:
                H = Ji.T @ Jj
                rows.append(np.repeat(6 * i + blk, 6)); cols.append(np.tile(6 * j + blk, 6)); vals.append(H.ravel())
    H = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(6 * P, 6 * P))
    delta = spla.spsolve(H, -g).reshape(P, 6)
    return retract(poses, delta), float(np.abs(delta).max()), cost


def hessian(factors, keys, poses, damp_keys=()):
    """Gauss-Newton Hessian J^T J (robust factors at their current Cauchy weights) as a dense matrix: small graphs only."""
    r, Ja, Jb, ia, ib, cost = linearize(factors, keys, poses)
    P = len(keys)
    H = np.zeros((6 * P, 6 * P))
    index = {int(k): i for i, k in enumerate(keys)}
    for dk in damp_keys:
        i = index[int(dk)]
        H[6 * i + np.arange(6), 6 * i + np.arange(6)] += np.array([1.0] * 3 + [4.0] * 3)
    for f in range(len(factors)):
        nodes = [(ib[f], Jb[f])] + ([(ia[f], Ja[f])] if ia[f] >= 0 else [])
        for (i, Ji) in nodes:
            for (j, Jj) in nodes:
                H[6 * i:6 * i + 6, 6 * j:6 * j + 6] += Ji.T @ Jj
    return H


def marginals(factors, keys, poses, query_keys, damp_keys=()):
    """gtsam::Marginals(graph, values).marginalCovariance(key) (reference laser_slam/src/laser_track.cpp:421-429):
    the 6x6 diagonal blocks of the inverse Hessian at `poses`, tangent order [translation; rotation]."""
    C = np.linalg.inv(hessian(factors, keys, poses, damp_keys))
    index = {int(k): i for i, k in enumerate(keys)}
    return np.stack([C[6 * index[int(k)]:6 * index[int(k)] + 6, 6 * index[int(k)]:6 * index[int(k)] + 6] for k in query_keys])


def optimize(factors, keys, poses, iters=3, tol=0.0, damp_keys=()):
    """`iters` Gauss-Newton iterations (3 = one IncrementalEstimator::estimate call).  Returns poses, history."""
    poses = np.asarray(poses, np.float64).copy()
    hist = []
    for _ in range(iters):
        poses, dmax, cost = gauss_newton_step(factors, keys, poses, damp_keys)
        hist.append((dmax, cost))
        if dmax < tol:
            break
    return poses, hist


# ---------------------------------------------------------------- synthetic config 4
def make_config4(n_poses=5000, n_lc=200, seed=4, outlier_frac=0.10, lap=1000):
    """Closed loop driven `n_poses/lap` times; 1 prior + odometry + ICP (Cauchy) between consecutive poses +
    `n_lc` loop closures between poses of different laps that are <= 5 m apart (10% gross outliers, Cauchy)."""
    rng = np.random.default_rng(seed)
    radius = 0.8 * lap / (2 * np.pi)
    ang = 2 * np.pi * np.arange(n_poses) / lap
    truth = np.zeros((n_poses, 4, 4))
    truth[:, 3, 3] = 1
    c, s = np.cos(ang + np.pi / 2), np.sin(ang + np.pi / 2)
    truth[:, 0, 0] = c; truth[:, 0, 1] = -s; truth[:, 1, 0] = s; truth[:, 1, 1] = c; truth[:, 2, 2] = 1
    truth[:, 0, 3] = radius * np.cos(ang); truth[:, 1, 3] = radius * np.sin(ang)
    truth[:, 2, 3] = 0.02 * np.sin(ang * 7)
    tp = se3_from_matrix(truth)

    def noisy(T, st, sr):
        d = np.concatenate([rng.normal(scale=st, size=3), rng.normal(scale=sr, size=3)])
        N = np.eye(4); N[:3, :3] = so3_exp(d[3:]); N[:3, 3] = d[:3]
        return T @ N

    sig = np.array([0.005] * 3 + [0.0015] * 3)
    keys = np.arange(100, 100 + n_poses, dtype=np.uint64)
    factors = [make_factor(PRIOR, keys[0], keys[0], tp[0], [1e-7] * 6)]
    odo = np.zeros((n_poses, 4, 4)); odo[0] = truth[0]
    for k in range(1, n_poses):
        rel = np.linalg.inv(truth[k - 1]) @ truth[k]
        m_odo = noisy(rel, 0.03, np.deg2rad(0.2))
        m_icp = noisy(rel, 0.004, 0.001)
        odo[k] = odo[k - 1] @ m_odo
        factors.append(make_factor(BETWEEN, keys[k - 1], keys[k], se3_from_matrix(m_odo), sig, robust=0))
        factors.append(make_factor(BETWEEN, keys[k - 1], keys[k], se3_from_matrix(m_icp), sig, robust=1))
    n_added = 0
    while n_added < n_lc:
        a = int(rng.integers(0, n_poses - lap))
        b = a + lap * int(rng.integers(1, (n_poses - a - 1) // lap + 1)) + int(rng.integers(-5, 6))
        if b <= a or b >= n_poses or np.linalg.norm(truth[a, :3, 3] - truth[b, :3, 3]) > 5.0:
            continue
        rel = np.linalg.inv(truth[a]) @ truth[b]
        if rng.random() < outlier_frac:
            rel = noisy(rel, 2.0, 0.3)
        else:
            rel = noisy(rel, 0.004, 0.001)
        factors.append(make_factor(BETWEEN, keys[a], keys[b], se3_from_matrix(rel), sig, robust=1))
        n_added += 1
    return keys, se3_from_matrix(odo), factors, tp
