import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import *
from mfr_b200 import pose_solver as ps
G = np.load(GOLDEN + "/pose_solver_reference.npz")
for hyp, seed in ((2048, 0x5EED), (2048, 7), (8192, 0x5EED)):
    cfg = {"EMAT_RANSAC": {"PIX_THRESHOLD": 2.0, "SCALE_THRESHOLD": 0.1, "CONFIDENCE": 0.9999},
           "PNP": {"RANSAC_ITER": 1000, "REPROJECTION_INLIER_THRESHOLD": 3, "CONFIDENCE": 0.9999},
           "GPU_RANSAC": {"NUM_HYPOTHESES": hyp, "SEED": seed}}
    pnp = ps.PnPSolver(cfg); em = ps.EssentialMatrixSolver(cfg)
    for ci in (0, 1, 2, 3, 5):
        c = pose_case(ci)
        d = {k: c[k] for k in ("K_color0", "K_color1", "depth0", "depth1")}
        R, t, n = pnp.estimate_pose(c["kpts0"], c["kpts1"], d)
        Re, te, ne = em.estimate_pose(c["kpts0"], c["kpts1"], d)
        print(f"hyp {hyp} seed {seed} case {ci}: PNP inl {n} (ref {int(G[f'c{ci}_PNP_inliers'])}) R vs ref {rot_angle(R, G[f'c{ci}_PNP_R']):.4f} vs gt {rot_angle(R, c['R_gt']):.4f} t vs gt {np.linalg.norm(t.ravel()-c['t_gt']):.4f} | "
              f"EMAT inl {ne} (ref {int(G[f'c{ci}_EssentialMatrix_inliers'])}) R vs gt {rot_angle(Re, c['R_gt']):.4f} (ref {rot_angle(G[f'c{ci}_EssentialMatrix_R'], c['R_gt']):.4f}) tdir vs gt {vec_angle(te, c['t_gt']):.4f} (ref {vec_angle(G[f'c{ci}_EssentialMatrix_t'], c['t_gt']):.4f})")
