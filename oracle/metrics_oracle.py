"""ORACLE (test infrastructure only — never imported by the product path).

numpy restatement of the Map-free leaderboard metrics: benchmark/metrics.py:40-67 (MetricManager), benchmark/utils.py:95-182
(quat_angle_error sine variant, precision_recall), benchmark/reprojection.py:7-86 (project, the virtual-point grid,
reprojection_error), benchmark/mapfree.py:76-117 (aggregate_results), benchmark/config.py. The reference imports
transforms3d (absent here) for four quaternion helpers; they are restated below from their published definitions
(w, x, y, z order; Hamilton product).

Pinned by the reference's OWN tests for this code, benchmark/test_metrics.py: its property tests and known-answer
vectors are replayed against this module in tests/test_metrics_oracle.py, and when /root/reference is mounted the
reference module itself is imported (with these helpers standing in for transforms3d) and compared on random poses."""
import numpy as np

T_THRESHOLD, R_THRESHOLD, VCRE_THRESHOLD = 0.25, 5, 90        # benchmark/config.py


# ---- transforms3d.quaternions restated
def quat2mat(q):
    w, x, y, z = np.asarray(q, dtype=np.float64)
    Nq = w * w + x * x + y * y + z * z
    if Nq < np.finfo(np.float64).eps:
        return np.eye(3)
    s = 2.0 / Nq
    X, Y, Z = x * s, y * s, z * s
    wX, wY, wZ, xX, xY, xZ, yY, yZ, zZ = w * X, w * Y, w * Z, x * X, x * Y, x * Z, y * Y, y * Z, z * Z
    return np.array([[1.0 - (yY + zZ), xY - wZ, xZ + wY], [xY + wZ, 1.0 - (xX + zZ), yZ - wX], [xZ - wY, yZ + wX, 1.0 - (xX + yY)]])


def qmult(q1, q2):
    w1, x1, y1, z1 = q1
    w2, x2, y2, z2 = q2
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2, w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2])


def qconjugate(q):
    return np.array(q) * np.array([1.0, -1, -1, -1])


def qinverse(q):
    return qconjugate(q) / np.dot(q, q)


def rotate_vector(v, q):
    return qmult(q, qmult(np.r_[0.0, v], qconjugate(q)))[1:]


def axangle2quat(vector, theta, is_normalized=False):
    vector = np.asarray(vector, dtype=np.float64)
    if not is_normalized:
        vector = vector / np.sqrt(np.dot(vector, vector))
    return np.concatenate([[np.cos(theta / 2.0)], vector * np.sin(theta / 2.0)])


def euler2quat(ai, aj, ak):        # 'sxyz' (transforms3d default): q = qz(ak) * qy(aj) * qx(ai)
    qx = axangle2quat([1, 0, 0], ai, True); qy = axangle2quat([0, 1, 0], aj, True); qz = axangle2quat([0, 0, 1], ak, True)
    return qmult(qz, qmult(qy, qx))


# ---- benchmark/utils.py:95-129
def quat_angle_error(label, pred):
    q1 = np.asarray(pred, np.float64) / np.linalg.norm(pred)
    q2 = np.asarray(label, np.float64) / np.linalg.norm(label)
    sine = qmult(q1, qinverse(q2))
    return float(np.arcsin(np.linalg.norm(sine[1:])) * 114.59155902616465)


# ---- benchmark/reprojection.py
def project(pts, K, img_size=None):
    uv_h = (K @ pts[:, :3].T).T
    uv = uv_h[:, :2] / uv_h[:, -1:]
    if img_size is not None:
        uv[:, 0] = np.clip(uv[:, 0], 0, img_size[0])
        uv[:, 1] = np.clip(uv[:, 1], 0, img_size[1])
    return uv


def get_grid_multipleheight():
    x = (np.arange(0, 7) - 3.0) * 0.3
    y = (np.arange(0, 4) - 1.5) * 0.3
    z = np.arange(0, 7).astype(float) * 0.3 + 1.8
    xx, yy, zz = np.meshgrid(x, y, z)
    ones = np.ones(xx.size)
    return np.concatenate([c.reshape(-1, 1) for c in (xx, yy, zz, ones)], axis=-1)


def reprojection_error(q_est, t_est, q_gt, t_gt, K, W, H):
    eye = get_grid_multipleheight()
    uv_gt = project(eye, K, (W, H))
    est = np.eye(4); est[:3, :3] = quat2mat(q_est); est[:3, -1] = t_est
    gt = np.eye(4); gt[:3, :3] = quat2mat(q_gt); gt[:3, -1] = t_gt
    res = (np.linalg.inv(est) @ gt @ eye.T).T
    uv_pred = project(res, K, (W, H))
    return float(np.linalg.norm(uv_gt - uv_pred, ord=2, axis=1).mean())


def pose_metrics(q_gt, t_gt, q_est, t_est, K, W, H):
    """MetricManager.__call__ for one pose (benchmark/metrics.py:40-67)."""
    return {"trans_err": float(np.linalg.norm(np.asarray(t_est) - np.asarray(t_gt))), "rot_err": quat_angle_error(label=q_est, pred=q_gt),
            "reproj_err": reprojection_error(q_est, t_est, q_gt, t_gt, K, W, H)}


# ---- benchmark/utils.py:132-182
def precision_recall(inliers, tp, failures):
    inliers = np.array(inliers)
    sort_idx = np.argsort(inliers)[::-1]
    inliers = inliers[sort_idx]
    tp = np.array(tp).reshape(-1)[sort_idx]
    distinct_value_indices = np.where(np.diff(inliers))[0]
    threshold_idxs = np.r_[distinct_value_indices, inliers.size - 1]
    N = inliers.shape[0]
    rec = np.arange(N, dtype=np.float32) + 1
    cum_tp = np.cumsum(tp)
    prec = cum_tp[threshold_idxs] / rec[threshold_idxs]
    rec = rec[threshold_idxs] / (float(N) + float(failures))
    last_ind = rec.searchsorted(rec[-1])
    sl = slice(last_ind, None, -1)
    prec = np.r_[prec[sl], 1]
    rec = np.r_[rec[sl], 0]
    return prec, rec, float(np.abs(np.sum(np.diff(rec) * np.array(prec)[:-1])))


# ---- benchmark/mapfree.py:76-117
def aggregate_results(all_results, all_failures):
    med, allm = {}, {}
    for m in ("trans_err", "rot_err", "reproj_err", "confidence"):
        med[m] = float(np.mean([np.median(r[m]) for r in all_results.values()]))
        allm[m] = np.concatenate([np.asarray(r[m], np.float64).reshape(-1) for r in all_results.values()])
    acc_pose = (allm["trans_err"] < T_THRESHOLD) * (allm["rot_err"] < R_THRESHOLD)
    acc_vcre = allm["reproj_err"] < VCRE_THRESHOLD
    total = len(allm["trans_err"]) + all_failures
    return {"Average Median Translation Error": med["trans_err"], "Average Median Rotation Error": med["rot_err"],
            "Average Median Reprojection Error": med["reproj_err"],
            "Precision @ Pose Error < (25.0cm, 5deg)": float(np.sum(acc_pose) / total),
            "AUC @ Pose Error < (25.0cm, 5deg)": precision_recall(allm["confidence"], acc_pose, all_failures)[2],
            "Precision @ VCRE < 90px": float(np.sum(acc_vcre) / total),
            "AUC @ VCRE < 90px": precision_recall(allm["confidence"], acc_vcre, all_failures)[2],
            "Estimates for % of frames": len(allm["trans_err"]) / total}
