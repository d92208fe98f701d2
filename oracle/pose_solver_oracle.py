"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the reference's online estimator back end,
``lib/models/matching/pose_solver.py`` of nianticlabs/map-free-reloc, on plain numpy arrays.
The arithmetic of the reference lives in OpenCV (un-vendored wheel, pinned
``opencv-python==4.8.0.74`` in the reference's environment.yml:17; this image has 4.13.0) and is
called here through exactly the same cv2 entry points and arguments as the reference's call
sites. Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference arm
may import this module.

Parity status: the reference has no tests or golden vectors for this path, so
this restatement is pinned against the *reference module itself*, imported from
/root/reference in the build container by ``tests/golden/make_golden.py``; the resulting vectors
are committed under ``tests/golden/`` and ``tests/test_oracle_golden.py`` replays them.
"""
import numpy as np
import cv2 as cv


def backproject_3d(uv, depth, K):
    """pose_solver.py:6-17. ``K`` arrives as a float32 tensor in the reference (model.py feeds
    data['K_color0'].squeeze(0)); np.linalg.inv therefore runs in float32 and the product is
    promoted to float64 by the integer/float64 ``uv1``."""
    uv1 = np.concatenate([uv, np.ones((uv.shape[0], 1))], axis=1)
    xyz = depth.reshape(-1, 1) * (np.linalg.inv(np.asarray(K)) @ uv1.T).T
    return xyz


def normalise_kpts(kpts0, kpts1, K0, K1):
    """pose_solver.py:39-40."""
    k0 = (kpts0 - K0[[0, 1], [2, 2]][None]) / K0[[0, 1], [0, 1]][None]
    k1 = (kpts1 - K1[[0, 1], [2, 2]][None]) / K1[[0, 1], [0, 1]][None]
    return k0, k1


def ransac_threshold(pix_threshold, K0, K1):
    """pose_solver.py:43 (mean of K0.fx, K1.fy, K0.fy, K1.fx)."""
    return pix_threshold / np.mean([K0[0, 0], K1[1, 1], K0[1, 1], K1[0, 0]])


def essential_matrix_solver(kpts0, kpts1, K0, K1, pix_threshold, confidence):
    """EssentialMatrixSolver.estimate_pose, pose_solver.py:29-61.
    Returns (R[3,3], t[3] or [3,1] NaN, n_inliers, mask) where ``mask`` is what the reference keeps
    in ``self.mask`` *after* cv.recoverPose mutated it in place (cheirality-filtered)."""
    R = np.full((3, 3), np.nan)
    t = np.full((3, 1), np.nan)
    if len(kpts0) < 5:
        return R, t, 0, None
    K0 = np.asarray(K0)
    K1 = np.asarray(K1)
    k0, k1 = normalise_kpts(kpts0, kpts1, K0, K1)
    thr = ransac_threshold(pix_threshold, K0, K1)
    E, mask = cv.findEssentialMat(k0, k1, np.eye(3), threshold=thr, prob=confidence,
                                  method=cv.USAC_MAGSAC)
    if E is None:
        return R, t, 0, mask
    best_num_inliers = 0
    ret = (R, t, 0)
    for _E in np.split(E, len(E) / 3):
        n, R, t, _ = cv.recoverPose(_E, k0, k1, np.eye(3), 1e9, mask=mask)
        if n > best_num_inliers:
            best_num_inliers = n
            ret = (R, t[:, 0], n)
    return ret[0], ret[1], ret[2], mask


def sampson_residuals(R, t, k0, k1):
    """Signed Sampson distance of normalised correspondences under E = [t]x R:
    r = x1^T E x0 / sqrt((E x0)_0^2 + (E x0)_1^2 + (E^T x1)_0^2 + (E^T x1)_1^2); OpenCV's inlier test is r^2 < thr^2."""
    t = np.asarray(t, dtype=np.float64).ravel()
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    E = tx @ np.asarray(R, dtype=np.float64)
    x0 = np.concatenate([k0, np.ones((len(k0), 1))], 1)
    x1 = np.concatenate([k1, np.ones((len(k1), 1))], 1)
    Ex0 = x0 @ E.T
    Etx1 = x1 @ E
    num = np.sum(x1 * Ex0, axis=1)
    den = np.sqrt(Ex0[:, 0] ** 2 + Ex0[:, 1] ** 2 + Etx1[:, 0] ** 2 + Etx1[:, 1] ** 2)
    return num / den


def msac_cost(R, t, kpts0, kpts1, K0, K1, pix_threshold):
    """sum min(Sampson^2, thr^2) / thr^2 over all correspondences: the objective the LO stages minimise."""
    K0 = np.asarray(K0, dtype=np.float64); K1 = np.asarray(K1, dtype=np.float64)
    k0, k1 = normalise_kpts(np.asarray(kpts0, np.float64), np.asarray(kpts1, np.float64), K0, K1)
    thr = ransac_threshold(pix_threshold, K0, K1)
    r2 = sampson_residuals(R, t, k0, k1) ** 2
    return float(np.minimum(r2, thr * thr).sum() / (thr * thr))


def polish_essential_fixed_point(R, t, kpts0, kpts1, K0, K1, pix_threshold, max_rounds=50):
    """Extension of ``polish_essential``: alternate (inlier set = Sampson^2 < thr^2 under the current model, the test that
    produces OpenCV's mask) and (Sampson least squares on that set) from the reference's output until the set stops
    changing. The fixed point is the model that is the least-squares optimum of its OWN inlier set, i.e. a local minimum
    of the MSAC cost sum min(r^2, thr^2); which one is reached is decided by the reference's output (the basin it lies
    in). An independent RANSAC whose final stage minimises the same cost (mfr_emat_ransac with lo_enabled = 3) must land
    on the same point whenever it ends up in the same basin. Returns (R, t, mask, rounds)."""
    K0 = np.asarray(K0, dtype=np.float64); K1 = np.asarray(K1, dtype=np.float64)
    k0, k1 = normalise_kpts(np.asarray(kpts0, np.float64), np.asarray(kpts1, np.float64), K0, K1)
    thr = ransac_threshold(pix_threshold, K0, K1)
    Rc, tc = np.asarray(R, np.float64), np.asarray(t, np.float64).ravel()
    prev = None
    for rounds in range(1, max_rounds + 1):
        m = sampson_residuals(Rc, tc, k0, k1) ** 2 < thr * thr
        if prev is not None and np.array_equal(m, prev):
            break
        prev = m
        if m.sum() < 5:
            break
        Rc, tc = polish_essential(Rc, tc, m.astype(np.uint8), kpts0, kpts1, K0, K1)
    return Rc, tc, prev, rounds


def polish_essential(R, t, mask, kpts0, kpts1, K0, K1):
    """The "polished oracle" of SURVEY.md §7 hard part 1(b): the reference solver's output (R, unit t, cheirality
    mask of pose_solver.py:46-60) -> deterministic Levenberg-Marquardt on the Sampson residuals of the reference's
    OWN inlier set (fixed), parameters = rotation vector increment (left-multiplied) + 2 tangent coordinates of t on
    the unit sphere. cv.findEssentialMat(USAC_MAGSAC) does not return the least-squares optimum of its inlier set;
    this does, and an independent solver can be compared with it at 1e-3 rad."""
    from scipy.optimize import least_squares
    K0 = np.asarray(K0, dtype=np.float64); K1 = np.asarray(K1, dtype=np.float64)
    k0, k1 = normalise_kpts(np.asarray(kpts0, np.float64), np.asarray(kpts1, np.float64), K0, K1)
    m = np.asarray(mask).ravel() == 1
    k0, k1 = k0[m], k1[m]
    R0 = np.asarray(R, dtype=np.float64)
    t0 = np.asarray(t, dtype=np.float64).ravel()
    t0 = t0 / np.linalg.norm(t0)
    a = np.eye(3)[np.argmin(np.abs(t0))]
    b1 = np.cross(t0, a); b1 /= np.linalg.norm(b1)
    b2 = np.cross(t0, b1)

    def unpack(p):
        dR, _ = cv.Rodrigues(p[:3].reshape(3, 1))
        tt = t0 + p[3] * b1 + p[4] * b2
        return dR @ R0, tt / np.linalg.norm(tt)

    def fun(p):
        Rp, tp = unpack(p)
        return sampson_residuals(Rp, tp, k0, k1)

    sol = least_squares(fun, np.zeros(5), method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=2000)
    return unpack(sol.x)


def scale_from_depth(R, t, mask, kpts0, kpts1, depth0, depth1, K0, K1, scale_threshold):
    """EssentialMatrixMetricSolver.estimate_pose, pose_solver.py:132-172, the part after the
    scale-free solve. depth0/depth1 are [H,W] float32 (the reference indexes data['depth0'][0])."""
    m = mask.ravel() == 1
    ik0 = np.int32(kpts0[m])
    ik1 = np.int32(kpts1[m])
    d0 = depth0[ik0[:, 1], ik0[:, 0]]
    d1 = depth1[ik1[:, 1], ik1[:, 0]]
    valid = (d0 > 0) * (d1 > 0)
    if valid.sum() < 1:
        return np.full((3, 3), np.nan), np.full((3, 1), np.nan), 0
    xyz0 = backproject_3d(ik0[valid], d0[valid], K0)
    xyz1 = backproject_3d(ik1[valid], d1[valid], K1)
    xyz0 = (R @ xyz0.T).T
    scale = np.dot(xyz1 - xyz0, t.reshape(3, 1))
    best_inliers = 0
    best_scale = None
    for scale_hyp in scale:
        inliers_hyp = (np.abs(scale - scale_hyp) < scale_threshold).sum().item()
        if inliers_hyp > best_inliers:
            best_scale = scale_hyp
            best_inliers = inliers_hyp
    t_metric = (best_scale * t).reshape(3, 1)
    return R, t_metric, best_inliers


def essential_matrix_metric_solver(kpts0, kpts1, depth0, depth1, K0, K1, pix_threshold,
                                   confidence, scale_threshold):
    """EssentialMatrixMetricSolver.estimate_pose, pose_solver.py:125-172."""
    R, t, inliers, mask = essential_matrix_solver(kpts0, kpts1, K0, K1, pix_threshold, confidence)
    if inliers == 0:
        return R, t, inliers
    return scale_from_depth(R, t, mask, kpts0, kpts1, depth0, depth1, K0, K1, scale_threshold)


def pnp_solver(pts0, pts1, depth0, K0, K1, ransac_iter, reproj_threshold, confidence,
               return_inliers=False):
    """PnPSolver.estimate_pose, pose_solver.py:184-235."""
    nan = (np.full((3, 3), np.nan), np.full((3, 1), np.nan), 0)
    pts0 = np.int32(pts0)
    if len(pts0) < 4:
        return nan + (None,) if return_inliers else nan
    depth_pts0 = depth0[pts0[:, 1], pts0[:, 0]]
    valid = depth_pts0 > depth0.min()
    if valid.sum() < 4:
        return nan + (None,) if return_inliers else nan
    pts0 = pts0[valid]
    pts1 = pts1[valid]
    depth_pts0 = depth_pts0[valid]
    xyz_0 = backproject_3d(pts0, depth_pts0, K0)
    K1 = np.asarray(K1)
    succ, rvec, tvec, inliers = cv.solvePnPRansac(
        xyz_0, pts1, K1, None, iterationsCount=ransac_iter,
        reprojectionError=reproj_threshold, confidence=confidence, flags=cv.SOLVEPNP_P3P)
    if succ and len(inliers) >= 6:
        succ, rvec, tvec, _ = cv.solvePnPGeneric(
            xyz_0[inliers], pts1[inliers], K1, None, useExtrinsicGuess=True, rvec=rvec, tvec=tvec,
            flags=cv.SOLVEPNP_ITERATIVE)
        rvec = rvec[0]
        tvec = tvec[0]
    if succ:
        if np.linalg.norm(tvec) > 1000:
            succ = False
    if succ:
        R, _ = cv.Rodrigues(rvec)
        t = tvec.reshape(3, 1)
    else:
        R = np.full((3, 3), np.nan)
        t = np.full((3, 1), np.nan)
        inliers = []
    if return_inliers:
        idx = np.where(valid)[0][np.asarray(inliers).ravel()] if len(inliers) else np.zeros(0, int)
        return R, t, len(inliers), idx
    return R, t, len(inliers)


def kabsch(P, Q):
    """Least-squares rigid transform Q ~ R P + t (lib/utils/solver.py:4-37 states the same maths)."""
    pb, qb = P.mean(0), Q.mean(0)
    Hm = (P - pb).T @ (Q - qb)
    U, S, Vt = np.linalg.svd(Hm)
    d = np.sign(np.linalg.det(Vt.T @ U.T))
    R = Vt.T @ np.diag([1.0, 1.0, d]) @ U.T
    return R, qb - R @ pb


def procrustes_solver(pts0, pts1, depth0, depth1, K0, K1, max_corr_dist, iters=4096, seed=0):
    """ProcrustesSolver.estimate_pose, pose_solver.py:247-320 with REFINE=False. PARITY UNPINNED for
    the RANSAC core: the reference calls Open3D 0.17 `registration_ransac_based_on_correspondence`,
    which cannot be installed here; it is replaced by the textbook algorithm it implements (3-point
    Kabsch hypotheses, inliers by 3-D distance < max_corr_dist, refit on the inliers). Everything
    around it (int truncation, depth > depth.min() validity, back-projection) follows the reference."""
    nan = (np.full((3, 3), np.nan), np.full((3, 1), np.nan), 0)
    pts0 = np.int32(pts0); pts1 = np.int32(pts1)
    if len(pts0) < 3:
        return nan
    d0 = depth0[pts0[:, 1], pts0[:, 0]]; d1 = depth1[pts1[:, 1], pts1[:, 0]]
    valid = (d0 > depth0.min()) * (d1 > depth1.min())
    if valid.sum() < 3:
        return nan
    X0 = backproject_3d(pts0[valid], d0[valid], K0); X1 = backproject_3d(pts1[valid], d1[valid], K1)
    rng = np.random.default_rng(seed)
    best = (-1, None)
    for _ in range(iters):
        idx = rng.choice(len(X0), 3, replace=False)
        R, t = kabsch(X0[idx], X1[idx])
        inl = np.linalg.norm(X0 @ R.T + t - X1, axis=1) < max_corr_dist
        if inl.sum() > best[0]:
            best = (int(inl.sum()), inl)
    inl = best[1]
    for _ in range(2):
        if inl.sum() < 3:
            return nan
        R, t = kabsch(X0[inl], X1[inl])
        inl = np.linalg.norm(X0 @ R.T + t - X1, axis=1) < max_corr_dist
    return R, t.reshape(3, 1), int(inl.sum())


def icp_point_to_point(src, dst, T_init, max_corr_dist, max_iter=30, rel_fitness=1e-4, rel_rmse=1e-4):
    """Open3D 0.17 registration_icp with TransformationEstimationPointToPoint, restated (PARITY UNPINNED: Open3D cannot
    be installed here; this follows its published algorithm — evaluate, then per iteration Kabsch on the current
    correspondences, re-evaluate, stop when fitness and inlier rmse both change by less than the relative criteria):
    nearest neighbour of every transformed source point in `dst` within max_corr_dist (exact, k-d tree),
    fitness = #correspondences / #source, inlier_rmse = sqrt(mean squared correspondence distance).
    Returns (T [4,4], fitness, inlier_rmse)."""
    from scipy.spatial import cKDTree
    tree = cKDTree(dst)
    T = np.array(T_init, dtype=np.float64)
    P = src @ T[:3, :3].T + T[:3, 3]

    def evaluate(P):
        d, j = tree.query(P, k=1, distance_upper_bound=max_corr_dist)
        ok = np.isfinite(d)
        n = int(ok.sum())
        return ok, j, (n / len(P) if len(P) else 0.0), (float(np.sqrt(np.mean(d[ok] ** 2))) if n else 0.0)

    ok, j, fit, rmse = evaluate(P)
    for _ in range(max_iter):
        if ok.sum() < 3:
            break
        Ru, tu = kabsch(P[ok], dst[j[ok]])
        U = np.eye(4); U[:3, :3] = Ru; U[:3, 3] = tu
        T = U @ T
        P = P @ Ru.T + tu
        ok, j, fit2, rmse2 = evaluate(P)
        done = abs(fit - fit2) < rel_fitness and abs(rmse - rmse2) < rel_rmse
        fit, rmse = fit2, rmse2
        if done:
            break
    return T, fit, rmse


def procrustes_icp_solver(pts0, pts1, depth0, depth1, K0, K1, max_corr_dist, iters=4096, seed=0):
    """ProcrustesSolver.estimate_pose with REFINE=True (pose_solver.py:289-319): the RANSAC pose, then ICP over the
    full clouds of both depth maps (every pixel with depth > 0), inliers = int(fitness * #target points)."""
    R, t, n = procrustes_solver(pts0, pts1, depth0, depth1, K0, K1, max_corr_dist, iters, seed)
    if n == 0:
        return R, t, n
    vv, uu = np.mgrid[0:depth0.shape[0], 0:depth1.shape[1]]
    uv = np.concatenate([uu.reshape(-1, 1), vv.reshape(-1, 1)], axis=1)
    v0 = depth0.reshape(-1) > 0
    X0 = backproject_3d(uv[v0], depth0.reshape(-1)[v0], K0)
    v1 = depth1.reshape(-1) > 0
    X1 = backproject_3d(uv[v1], depth1.reshape(-1)[v1], K1)
    T0 = np.eye(4); T0[:3, :3] = R; T0[:3, 3] = np.ravel(t)
    T, fit, _ = icp_point_to_point(X0, X1, T0, max_corr_dist)
    return T[:3, :3], T[:3, 3].reshape(3, 1), int(fit * len(X1))


# --------------------------------------------------------------------------------------------
# Oracle extensions (not in the reference): the "polished oracle" of SURVEY.md §7 hard part 1
# --------------------------------------------------------------------------------------------
def sampson_sq(E, k0, k1):
    """Squared Sampson distance of normalised correspondences under x1^T E x0 = 0 (the quantity
    OpenCV thresholds to build the findEssentialMat mask)."""
    x0 = np.concatenate([k0, np.ones((len(k0), 1))], 1)
    x1 = np.concatenate([k1, np.ones((len(k1), 1))], 1)
    Ex0 = x0 @ E.T
    Etx1 = x1 @ E
    num = np.sum(x1 * Ex0, 1) ** 2
    den = Ex0[:, 0] ** 2 + Ex0[:, 1] ** 2 + Etx1[:, 0] ** 2 + Etx1[:, 1] ** 2
    return num / den


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def _expm_so3(w):
    th = np.linalg.norm(w)
    if th < 1e-14:
        return np.eye(3) + _skew(w)
    k = w / th
    Kx = _skew(k)
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def polish_pose_sampson(R, t, k0, k1, mask, iters=50):
    """Deterministic Gauss-Newton/LM on the Sampson error over a FIXED inlier mask, pose
    parameterised as (so(3) increment, tangent-plane increment of the unit translation).
    Returns the locally optimal (R, t_unit)."""
    from scipy.optimize import least_squares
    m = np.asarray(mask).ravel().astype(bool)
    a0, a1 = k0[m].astype(np.float64), k1[m].astype(np.float64)
    R0 = np.asarray(R, dtype=np.float64)
    t0 = np.asarray(t, dtype=np.float64).ravel()
    t0 = t0 / np.linalg.norm(t0)
    # tangent basis of the sphere at t0
    a = np.array([1.0, 0, 0]) if abs(t0[0]) < 0.9 else np.array([0, 1.0, 0])
    b1 = np.cross(t0, a)
    b1 /= np.linalg.norm(b1)
    b2 = np.cross(t0, b1)

    def unpack(p):
        Rn = _expm_so3(p[:3]) @ R0
        tn = t0 + p[3] * b1 + p[4] * b2
        tn = tn / np.linalg.norm(tn)
        return Rn, tn

    def resid(p):
        Rn, tn = unpack(p)
        E = _skew(tn) @ Rn
        x0 = np.concatenate([a0, np.ones((len(a0), 1))], 1)
        x1 = np.concatenate([a1, np.ones((len(a1), 1))], 1)
        Ex0 = x0 @ E.T
        Etx1 = x1 @ E
        num = np.sum(x1 * Ex0, 1)
        den = np.sqrt(Ex0[:, 0] ** 2 + Ex0[:, 1] ** 2 + Etx1[:, 0] ** 2 + Etx1[:, 1] ** 2)
        return num / den

    sol = least_squares(resid, np.zeros(5), method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15,
                        max_nfev=200 * iters)
    return unpack(sol.x)


def rotation_angle(Ra, Rb):
    """angle(Ra Rb^T) in radians (lib/utils/metrics.py:24-28 formula)."""
    c = (np.trace(Ra @ Rb.T) - 1.0) / 2.0
    return float(np.arccos(np.clip(c, -1.0, 1.0)))


def vector_angle(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    c = np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))
    return float(np.arccos(np.clip(c, -1.0, 1.0)))
