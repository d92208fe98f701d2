"""E-mat parity protocol of SURVEY.md §7 hard part 1, measured: for seeded synthetic correspondence sets,
(a) inlier-mask IoU vs the OpenCV oracle, (b) R / t-direction distance to the POLISHED oracle (reference output ->
Sampson LM on its own mask), (c) raw-oracle gap and errors vs ground truth, for both final-refit modes of the GPU
solver. Run on the GPU box: python scripts/emat_protocol.py > gpurun_out/emat_protocol.txt"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfr_b200  # noqa: E402,F401
from helpers import rot_angle, vec_angle  # noqa: E402
from mfr_b200 import pose_solver as ps, synth  # noqa: E402
from oracle import pose_solver_oracle as po  # noqa: E402

CASES = [(41, 1000, 0.6, 0.3), (42, 2000, 0.5, 0.5), (43, 800, 0.7, 0.2), (44, 3000, 0.4, 0.5), (45, 5000, 0.5, 0.5),
         (46, 1500, 0.8, 0.1), (47, 600, 0.5, 1.0), (51, 600, 0.8, 0.0)]

if __name__ == "__main__":
    base = {"EMAT_RANSAC": {"PIX_THRESHOLD": 2.0, "SCALE_THRESHOLD": 0.1, "CONFIDENCE": 0.9999}}
    solvers = {r: ps.EssentialMatrixMetricSolver(dict(base, GPU_RANSAC={"FINAL_REFIT": r})) for r in ("lsq", "tukey")}
    for seed, n, ratio, noise in CASES:
        c = synth.make_correspondences(seed, n=n, inlier_ratio=ratio, noise_px=noise)
        Ro, to, no, mo = po.essential_matrix_solver(c["kpts0"], c["kpts1"], c["K_color0"], c["K_color1"], 2.0, 0.9999)
        Rp, tp = po.polish_essential(Ro, to, mo, c["kpts0"], c["kpts1"], c["K_color0"], c["K_color1"])
        Rf, tf, mf, rounds = po.polish_essential_fixed_point(Ro, to, c["kpts0"], c["kpts1"], c["K_color0"], c["K_color1"], 2.0)
        data = {k: c[k] for k in ("K_color0", "K_color1", "depth0", "depth1")}
        for name, s in solvers.items():
            R, t, ninl = ps.EssentialMatrixSolver.estimate_pose(s, c["kpts0"], c["kpts1"], data)
            m = s.mask.ravel().astype(bool); mo_ = mo.ravel().astype(bool)
            row = {"seed": seed, "n": n, "inlier_ratio": ratio, "noise_px": noise, "refit": name, "hyp": s.last_num_hyp,
                   "mask_iou": round(float((m & mo_).sum() / max(1, (m | mo_).sum())), 4), "inliers": int(ninl), "cv_inliers": int(no),
                   "R_vs_polished_rad": rot_angle(R, Rp), "t_vs_polished_rad": vec_angle(t, tp),
                   "R_vs_fixed_point_rad": rot_angle(R, Rf), "t_vs_fixed_point_rad": vec_angle(t, tf), "fixed_point_rounds": rounds,
                   "fixed_point_R_err_gt": rot_angle(Rf, c["R_gt"]), "fixed_point_t_err_gt": vec_angle(tf, c["t_gt"]),
                   "msac_cost": po.msac_cost(R, t, c["kpts0"], c["kpts1"], c["K_color0"], c["K_color1"], 2.0),
                   "fixed_point_msac_cost": po.msac_cost(Rf, tf, c["kpts0"], c["kpts1"], c["K_color0"], c["K_color1"], 2.0),
                   "cv_msac_cost": po.msac_cost(Ro, to, c["kpts0"], c["kpts1"], c["K_color0"], c["K_color1"], 2.0),
                   "R_vs_raw_rad": rot_angle(R, Ro), "t_vs_raw_rad": vec_angle(t, to),
                   "R_err_gt": rot_angle(R, c["R_gt"]), "t_err_gt": vec_angle(t, c["t_gt"]),
                   "cv_R_err_gt": rot_angle(Ro, c["R_gt"]), "cv_t_err_gt": vec_angle(to, c["t_gt"]),
                   "polished_R_err_gt": rot_angle(Rp, c["R_gt"]), "polished_t_err_gt": vec_angle(tp, c["t_gt"])}
            print(json.dumps({k: (float(f"{v:.3e}") if isinstance(v, float) else v) for k, v in row.items()}), flush=True)
