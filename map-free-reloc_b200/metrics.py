"""Leaderboard metrics of the Map-free benchmark on the device (reference: benchmark/metrics.py, benchmark/utils.py:95-182,
benchmark/reprojection.py, benchmark/mapfree.py:76-117, benchmark/config.py). The per-pose metrics run in a CUDA kernel
(libmfr_b200.so: mfr_pose_metrics, one warp per pose); the aggregation — per-scene medians, precision and the
precision/recall AUC over the confidence ranking — is a handful of torch sort / cumsum calls on the same device.
No CPU fallback: the functions need CUDA tensors."""
import torch

from . import lib as _lib

T_THRESHOLD, R_THRESHOLD, VCRE_THRESHOLD = 0.25, 5.0, 90.0        # benchmark/config.py


def pose_metrics(q_gt, t_gt, q_est, t_est, K, W, H):
    """q_* [n,4] (w,x,y,z), t_* [n,3], K [n,3,3]: CUDA tensors (any float dtype; computed in float64 like the reference).
    Returns {'trans_err' [m], 'rot_err' [deg], 'reproj_err' [px]} float64 [n] on the same device."""
    if not q_gt.is_cuda:
        raise _lib.MfrError("mfr_b200.metrics needs CUDA tensors (no CPU fallback)")
    l = _lib.load()
    n = q_gt.shape[0]
    a = [x.to(torch.float64).contiguous() for x in (q_gt.reshape(n, 4), t_gt.reshape(n, 3), q_est.reshape(n, 4), t_est.reshape(n, 3),
                                                    K.reshape(n, 9))]
    out = [torch.empty(n, dtype=torch.float64, device=q_gt.device) for _ in range(3)]
    _lib.check(l.mfr_pose_metrics(*[_lib.ptr(x) for x in a], int(W), int(H), n, *[_lib.ptr(x) for x in out], _lib.cur_stream()),
               "mfr_pose_metrics")
    return {"trans_err": out[0], "rot_err": out[1], "reproj_err": out[2]}


def precision_recall(inliers, tp, failures):
    """benchmark/utils.py:132-182: precision / recall over the confidence (inlier count) ranking, evaluated where the
    confidence changes, and the average precision (AUC). Tensors on any device; returns (prec, rec, ap)."""
    inliers = inliers.reshape(-1).to(torch.float64)
    tp = tp.reshape(-1).to(torch.float64)
    N = inliers.numel()
    if N == 0:
        z = torch.zeros(0, dtype=torch.float64, device=inliers.device)
        return z, z, 0.0
    order = torch.argsort(inliers, descending=True, stable=True)
    inl, tp = inliers[order], tp[order]
    distinct = torch.nonzero(inl[1:] != inl[:-1]).reshape(-1)
    thr = torch.cat([distinct, torch.tensor([N - 1], device=inl.device)])
    rec_all = torch.arange(1, N + 1, dtype=torch.float32, device=inl.device)
    prec = (torch.cumsum(tp, 0)[thr] / rec_all[thr].to(torch.float64))
    rec = rec_all[thr] / torch.tensor(float(N) + float(failures), dtype=torch.float32, device=inl.device)   # float32, like numpy's
    last = int(torch.searchsorted(rec, rec[-1:].contiguous()).item())
    idx = torch.arange(last, -1, -1, device=inl.device)
    prec = torch.cat([prec[idx], torch.ones(1, dtype=torch.float64, device=inl.device)])
    rec = torch.cat([rec[idx], torch.zeros(1, dtype=torch.float32, device=inl.device)])
    ap = float(torch.abs(torch.sum((rec[1:] - rec[:-1]).to(torch.float64) * prec[:-1])).item())    # np.diff in float32
    return prec, rec.to(torch.float64), ap


def aggregate_results(all_results, all_failures):
    """benchmark/mapfree.py:76-117. ``all_results``: {scene: {'trans_err', 'rot_err', 'reproj_err', 'confidence': 1-D tensors}}."""
    names = ("trans_err", "rot_err", "reproj_err", "confidence")
    med = {m: torch.stack([torch.quantile(r[m].to(torch.float64), 0.5, interpolation="midpoint") for r in all_results.values()]).mean().item()
           for m in names}                          # np.median = midpoint of the two central values
    allm = {m: torch.cat([r[m].to(torch.float64).reshape(-1) for r in all_results.values()]) for m in names}
    accepted_pose = (allm["trans_err"] < T_THRESHOLD) & (allm["rot_err"] < R_THRESHOLD)
    accepted_vcre = allm["reproj_err"] < VCRE_THRESHOLD
    total = allm["trans_err"].numel() + all_failures
    _, _, auc_pose = precision_recall(allm["confidence"], accepted_pose, all_failures)
    _, _, auc_vcre = precision_recall(allm["confidence"], accepted_vcre, all_failures)
    return {
        "Average Median Translation Error": med["trans_err"],
        "Average Median Rotation Error": med["rot_err"],
        "Average Median Reprojection Error": med["reproj_err"],
        f"Precision @ Pose Error < ({T_THRESHOLD * 100}cm, {R_THRESHOLD:g}deg)": accepted_pose.sum().item() / total,
        f"AUC @ Pose Error < ({T_THRESHOLD * 100}cm, {R_THRESHOLD:g}deg)": auc_pose,
        f"Precision @ VCRE < {VCRE_THRESHOLD:g}px": accepted_vcre.sum().item() / total,
        f"AUC @ VCRE < {VCRE_THRESHOLD:g}px": auc_vcre,
        "Estimates for % of frames": allm["trans_err"].numel() / total,
    }
