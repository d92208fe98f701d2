"""BASELINE config C5: RANSAC hypothesis sweep 256-65536 hypotheses/pair at inlier ratios 5-50 %
(warp-per-hypothesis E-matrix kernel + LO + scale from depth), P pairs per call. Prints a table of device time
per pair, success rate (rot < 1 deg and t < 10 cm vs ground truth) and median errors."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mfr_b200
from mfr_b200 import pose_solver, synth

P = int(os.environ.get("P", 16)); N = int(os.environ.get("N", 2000))
be = pose_solver.PoseBackend(max_pairs=P, max_pts=N, max_hyp=65536)
print(f"# {P} pairs per call, {N} correspondences per pair, 0.5 px noise, 720x540, K f=600; device time by CUDA events (5 calls)")
print("inlier_ratio  hypotheses  ms_per_call  us_per_pair  success  med_rot_deg  med_t_m")
for ratio in (0.05, 0.1, 0.2, 0.5):
    cs = [synth.make_correspondences(5000 + i, n=N, inlier_ratio=ratio, noise_px=0.5) for i in range(P)]
    k0 = torch.from_numpy(np.concatenate([c["kpts0"] for c in cs])).cuda(); k1 = torch.from_numpy(np.concatenate([c["kpts1"] for c in cs])).cuda()
    off = torch.arange(P + 1, dtype=torch.int32, device="cuda") * N
    K = torch.from_numpy(np.stack([c["K_color0"] for c in cs]).astype(np.float32)).cuda().contiguous()
    d0 = torch.from_numpy(np.stack([c["depth0"] for c in cs]).astype(np.float32)).cuda(); d1 = torch.from_numpy(np.stack([c["depth1"] for c in cs]).astype(np.float32)).cuda()
    for hyp in (256, 1024, 4096, 16384, 65536):
        for _ in range(2):
            R, t, n = be.emat_metric(k0, k1, off, K, K, d0, d1, 2.0, 0.3, hyp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            R, t, n = be.emat_metric(k0, k1, off, K, K, d0, d1, 2.0, 0.3, hyp)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        Rn, tn = R.cpu().numpy().astype(np.float64), t.cpu().numpy().astype(np.float64)
        re, te = [], []
        for i, c in enumerate(cs):
            if not np.all(np.isfinite(Rn[i])):
                re.append(np.inf); te.append(np.inf); continue
            re.append(np.degrees(np.arccos(np.clip((np.trace(Rn[i].T @ c["R_gt"]) - 1) / 2, -1, 1)))); te.append(np.linalg.norm(tn[i] - c["t_gt"]))
        re, te = np.array(re), np.array(te)
        print(f"{ratio:12.2f}  {hyp:10d}  {ms:11.3f}  {ms * 1e3 / P:11.1f}  {np.mean((re < 1.0) & (te < 0.1)):7.2f}  {np.median(re):11.4f}  {np.median(te):7.4f}", flush=True)
