"""GPU pose back end vs the OpenCV oracle on synthetic correspondences (prints errors + timing)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mfr_b200
from mfr_b200 import synth, pose_solver as ps
from oracle import pose_solver_oracle as po

cfg = {"EMAT_RANSAC": {"PIX_THRESHOLD": 2.0, "SCALE_THRESHOLD": 0.1, "CONFIDENCE": 0.9999},
       "PNP": {"RANSAC_ITER": 1000, "REPROJECTION_INLIER_THRESHOLD": 3, "CONFIDENCE": 0.9999},
       "GPU_RANSAC": {"NUM_HYPOTHESES": int(os.environ.get("HYP", 2048))}}
em = ps.EssentialMatrixMetricSolver(cfg)
pnp = ps.PnPSolver(cfg)
rows = []
for seed, n, ratio, noise in [(1, 1000, 0.5, 0.5), (2, 2000, 0.3, 0.5), (3, 500, 0.7, 0.0), (4, 2000, 0.1, 0.5),
                              (5, 1000, 0.5, 1.0), (6, 300, 0.5, 0.1), (7, 4000, 0.5, 0.5), (8, 6000, 0.2, 0.5)]:
    c = synth.make_correspondences(seed, n=n, inlier_ratio=ratio, noise_px=noise)
    data = {k: c[k] for k in ("K_color0", "K_color1", "depth0", "depth1")}
    t0 = time.time()
    Ro, to, no, mo = po.essential_matrix_solver(c["kpts0"], c["kpts1"], c["K_color0"], c["K_color1"], 2.0, 0.9999)
    t_cv = time.time() - t0
    k0n, k1n = po.normalise_kpts(c["kpts0"], c["kpts1"], c["K_color0"], c["K_color1"])
    Rp, tp = po.polish_pose_sampson(Ro, to, k0n, k1n, mo)
    torch.cuda.synchronize(); t0 = time.time()
    R, t, ninl = ps.EssentialMatrixSolver.estimate_pose(em, c["kpts0"], c["kpts1"], data)
    torch.cuda.synchronize(); t_gpu = time.time() - t0
    m = em.mask.ravel().astype(bool); mo_b = mo.ravel().astype(bool)
    iou = (m & mo_b).sum() / max(1, (m | mo_b).sum())
    print(f"[emat n={n} in={ratio} noise={noise}] cv {t_cv*1e3:.1f}ms gpu {t_gpu*1e3:.1f}ms | inl cv {no} gpu {ninl} IoU {iou:.3f} | "
          f"R vs raw {po.rotation_angle(R, Ro):.2e} vs polished {po.rotation_angle(R, Rp):.2e} | t-dir vs raw {po.vector_angle(t, to):.2e} vs polished {po.vector_angle(t, tp):.2e} | "
          f"vs GT: gpu R {po.rotation_angle(R, c['R_gt']):.2e} t {po.vector_angle(t, c['t_gt']):.2e}; cv R {po.rotation_angle(Ro, c['R_gt']):.2e} t {po.vector_angle(to, c['t_gt']):.2e}")
    # metric
    Rm, tm, nm = em.estimate_pose(c["kpts0"], c["kpts1"], data)
    Rmo, tmo, nmo = po.essential_matrix_metric_solver(c["kpts0"], c["kpts1"], c["depth0"], c["depth1"], c["K_color0"], c["K_color1"], 2.0, 0.9999, 0.1)
    print(f"    metric: inl gpu {nm} cv {nmo} | |t-t_cv| {np.linalg.norm(tm.ravel()-tmo.ravel()):.2e} m | |t-t_gt| gpu {np.linalg.norm(tm.ravel()-c['t_gt']):.2e} cv {np.linalg.norm(tmo.ravel()-c['t_gt']):.2e}")
    # scale stage in isolation: feed the oracle's pose + mask
    k0, k1, off, K0, K1, _ = em._dev
    Rd = torch.tensor(Ro, dtype=torch.float64, device="cuda").reshape(1, 9).contiguous()
    td = torch.tensor(np.asarray(to).ravel(), dtype=torch.float64, device="cuda").reshape(1, 3).contiguous()
    mk = torch.tensor(mo.ravel().astype(np.uint8), device="cuda")
    d0 = torch.tensor(c["depth0"], device="cuda")[None]; d1 = torch.tensor(c["depth1"], device="cuda")[None]
    Rs, ts, ns = em.backend.scale(k0, k1, off, K0, K1, mk, d0, d1, 0.1, Rd, td)
    print(f"    scale stage (oracle pose in): inl {int(ns.item())} vs {nmo}; |dt| {np.linalg.norm(ts[0].cpu().numpy().astype(np.float64) - tmo.ravel()):.2e}")
    # pnp
    t0 = time.time()
    Rq, tq, nq, idx = po.pnp_solver(c["kpts0"], c["kpts1"], c["depth0"], c["K_color0"], c["K_color1"], 1000, 3, 0.9999, True)
    t_cv = time.time() - t0
    torch.cuda.synchronize(); t0 = time.time()
    Rg, tg, ng = pnp.estimate_pose(c["kpts0"], c["kpts1"], data)
    torch.cuda.synchronize(); t_gpu = time.time() - t0
    print(f"    pnp: cv {t_cv*1e3:.1f}ms gpu {t_gpu*1e3:.1f}ms inl cv {nq} gpu {ng} | R diff {po.rotation_angle(Rg, Rq):.2e} t diff {np.linalg.norm(tg.ravel()-tq.ravel()):.2e} | "
          f"vs GT gpu R {po.rotation_angle(Rg, c['R_gt']):.2e} t {np.linalg.norm(tg.ravel()-c['t_gt']):.2e}; cv R {po.rotation_angle(Rq, c['R_gt']):.2e} t {np.linalg.norm(tq.ravel()-c['t_gt']):.2e}")
# batched timing: 64 pairs
P = 64
cs = [synth.make_correspondences(100 + i, n=2000, inlier_ratio=0.3, noise_px=0.5) for i in range(8)]
k0 = torch.cat([torch.tensor(cs[i % 8]["kpts0"]) for i in range(P)]).cuda(); k1 = torch.cat([torch.tensor(cs[i % 8]["kpts1"]) for i in range(P)]).cuda()
off = torch.arange(0, P + 1, dtype=torch.int32).cuda() * 2000
K = torch.tensor(cs[0]["K_color0"]).reshape(1, 3, 3).repeat(P, 1, 1).cuda().contiguous()
d0 = torch.stack([torch.tensor(cs[i % 8]["depth0"]) for i in range(P)]).cuda(); d1 = torch.stack([torch.tensor(cs[i % 8]["depth1"]) for i in range(P)]).cuda()
be = ps.PoseBackend(max_pairs=P, max_hyp=8192)
for hyp in (256, 1024, 4096):
    for name, fn in (("emat_metric", lambda: be.emat_metric(k0, k1, off, K, K, d0, d1, 2.0, 0.1, hyp)), ("pnp", lambda: be.pnp(k0, k1, off, K, K, d0, 3.0, hyp))):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): out = fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"batched {name}: P={P} hyp={hyp}: {ms:.3f} ms/batch -> {P/ms*1e3:.0f} pairs/s")
