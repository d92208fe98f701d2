#!/usr/bin/env python
"""Benchmark of the feature-matching + scale-from-depth hot path (BASELINE.json metric:
image-pairs/sec @ 540x720).

Workload (BASELINE.json configs[1]): LoFTR coarse-to-fine matching + 5-point essential matrix with the
scale recovered from depth (the reference's config/matching/mapfree/loftr_emat_dptkitti.yaml:
EssentialMatrixMetric, 2 px / 0.1 / 0.9999), synthetic 540x720 gray+depth pairs, synthetic LoFTR weights
(the real checkpoints are not distributed with the reference). A "step" is one batch of B pairs through
matcher + solver on one GPU; with N GPUs every rank runs its own pairs (weak scaling), one NCCL gather
of the pose records at the end of the stream.

  python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU, torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W   # reference arm: CPU path on the host cores

Prints ONE JSON line (rank 0). `value`: inputs resident in HBM. `e2e`: host (pinned) buffers in, pose
records out, copies inside the timed region, through the public Python API.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H0, W0 = 720, 540            # dataset frame (config/mapfree.yaml HEIGHT/WIDTH)
PIX_THR, SCALE_THR = 2.0, 0.1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("MFR_BENCH_BATCH", 16)), help="pairs per step per GPU")
    ap.add_argument("--hyp", type=int, default=1024, help="5-point hypotheses per pair")
    ap.add_argument("--match-thr", type=float, default=0.2, help="LoFTR coarse confidence threshold (reference: 0.2)")
    ap.add_argument("--precision", default=os.environ.get("MFR_BENCH_PRECISION", "bf16"), choices=["tf32", "bf16"],
                    help="LoFTR backbone tensor-core input type (tf32 = parity mode; bf16 = bf16 storage + MMA in the ResNet-FPN)")
    ap.add_argument("--engines", type=int, default=int(os.environ.get("MFR_BENCH_ENGINES", 1)),
                    help="LoFTR engine instances that consecutive batches alternate between (streams overlap)")
    ap.add_argument("--pool", type=int, default=32, help="distinct synthetic pairs cycled through")
    ap.add_argument("--cpu-sample", type=int, default=4, help="pairs timed for the cpu_baseline leg")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------
def make_pool(n, seed0):
    from mfr_b200 import synth
    g0, g1, d0, d1, K, gt = [], [], [], [], [], []
    for i in range(n):
        p = synth.make_pair(seed0 + i)
        g0.append(synth.rgb_to_gray_u8(p["image0"]).astype(np.float32) / 255.0)
        g1.append(synth.rgb_to_gray_u8(p["image1"]).astype(np.float32) / 255.0)
        d0.append(p["depth0"]); d1.append(p["depth1"]); K.append(p["K_color0"])
        gt.append((p["R_gt"], p["t_gt"]))
    return (np.stack(g0), np.stack(g1), np.stack(d0), np.stack(d1), np.stack(K).astype(np.float32), gt)


def pose_errors(R, t, gt):
    """Rotation error (deg) and translation error (m) of one estimate against ground truth; NaN -> inf."""
    Rg, tg = gt
    if not (np.all(np.isfinite(R)) and np.all(np.isfinite(t))):
        return float("inf"), float("inf")
    c = np.clip((np.trace(np.asarray(R, np.float64).T @ Rg) - 1.0) / 2.0, -1.0, 1.0)
    return float(np.degrees(np.arccos(c))), float(np.linalg.norm(np.ravel(t).astype(np.float64) - np.ravel(tg)))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self._stop, self._t = gpu_index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


WEIGHTS = ("hand-constructed functional LoFTR weights (loftr.functional_state_dict: full-shape layers that implement a patch "
           "descriptor + normalised cross-correlation); the reference does not ship its checkpoints")
WORKLOAD = "LoFTR coarse-to-fine + 5-pt essential (EssentialMatrixMetric, scale from depth), 540x720 (720x544 padded), 1xB200 per rank"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured", d.get("bf16_tflops_sustained")
    return 6650.0, 1590.0, "fallback", 1400.0


def load_traffic(precision, batch):
    """DRAM bytes of the ResNet-FPN conv launches of one step, from the committed ncu capture of this
    workload (profiles/README.md); None when no capture exists for the configuration."""
    p = os.path.join(ROOT, "profiles", f"r01_conv_gemm_metrics_{precision}_B{batch}.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))["backbone"]
    return int(d["dram_read"] + d["dram_write"]), os.path.relpath(p, ROOT)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import mfr_b200  # noqa: F401
    from mfr_b200 import loftr, pose_solver

    rank, local_rank, world = dist_env()
    assert torch.cuda.is_available(), "bench.py needs CUDA devices (no CPU fallback in the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)
    HP, WP = H0 + H0 % 8, W0 + W0 % 8                       # matchers.py:41-46 -> 720 x 544

    eng = loftr.LoFTR(loftr.functional_state_dict(), batch=B, height=HP, width=WP, cfg={"thr": args.match_thr},
                      device=local_rank, graph=True, precision=args.precision)
    L = (HP // 8) * (WP // 8)
    be = pose_solver.PoseBackend(max_pairs=B, max_pts=L, max_hyp=max(args.hyp, 128), device=local_rank)
    g0, g1, d0, d1, Kc, gt = make_pool(args.pool, 1000 + 100 * rank)
    n_pool = args.pool

    def batch_idx(step):
        return [(step * B + b) % n_pool for b in range(B)]

    # pinned host copies (e2e) and device-resident copies (value)
    h = {k: torch.from_numpy(v).pin_memory() for k, v in (("g0", g0), ("g1", g1), ("d0", d0), ("d1", d1), ("K", Kc))}
    dres = {k: v.to(dev) for k, v in h.items()}
    from mfr_b200.pipeline import RelocPipeline
    outs = eng.device_outputs()
    # the repo's public streaming API: copy / matcher / RANSAC stages of consecutive batches overlap on three streams
    engines = [eng] + [loftr.LoFTR(loftr.functional_state_dict(), batch=B, height=HP, width=WP, cfg={"thr": args.match_thr},
                                   device=local_rank, graph=True, precision=args.precision) for _ in range(args.engines - 1)]
    pipe = RelocPipeline(engines, be, PIX_THR, SCALE_THR, args.hyp, depth_hw=(H0, W0))
    idx_cache = [torch.tensor(batch_idx(s), device=dev) for s in range(n_pool)]   # period divides n_pool

    def step_resident(step):
        """inputs already in HBM; returns the poses of the previous step's batch (one-deep pipeline)"""
        idx = idx_cache[step % n_pool]
        with torch.cuda.stream(engines[pipe.i % len(engines)].stream):
            a, b_, p0, p1, Kb = (dres["g0"][idx], dres["g1"][idx], dres["d0"][idx], dres["d1"][idx], dres["K"][idx].contiguous())
        return pipe.submit_device(a, b_, p0, p1, Kb)

    def step_e2e(step):
        """host (pinned) inputs -> device copies inside the step; the previous batch's poses are read back"""
        idx = batch_idx(step)
        return pipe.submit_host(*[[h[k][i] for i in idx] for k in ("g0", "g1", "d0", "d1", "K")])

    in_bytes = sum(int(np.prod(h[k].shape[1:])) * h[k].element_size() * B for k in h)
    out_bytes = int(pipe.slots[0]["rec_host"].numel() * 4)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, gather_at_end=False):
        for s in range(Wm):
            fn(s)
        pipe.drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(eng.stream):
            e0.record()
        for s in range(K):
            fn(Wm + s)
        last = pipe.drain()                          # the timed region ends when the last batch's poses are on the host
        if gather_at_end and world > 1:
            R, t, n = last
            with torch.cuda.stream(eng.stream):
                rec = torch.cat([R.reshape(B, 9), t.reshape(B, 3), n.reshape(B, 1).float()], 1).to(dev).contiguous()
                allr = [torch.empty_like(rec) for _ in range(world)]
                dist.all_gather(allr, rec)   # the path's only collective: 52-byte records per pair
        with torch.cuda.stream(eng.stream):
            eng.stream.wait_stream(pipe.s_solve)
            for e_ in engines[1:]:
                eng.stream.wait_stream(e_.stream)
            e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), last

    with ClockSampler(local_rank) as clk:
        ms_res, last = timed(step_resident, gather_at_end=True)
        clocks = clk.summary()
    ms_e2e, _ = timed(step_e2e)
    pairs = K * B * world
    value = pairs / (ms_res / 1e3)
    e2e_value = pairs / (ms_e2e / 1e3)
    with torch.cuda.stream(eng.stream):
        m_last = int(outs["m_total"][0].item())

    # ---- accuracy pass (untimed): every pair of the pool once through the host-buffer path
    acc_rows = {}
    n_acc = (n_pool + B - 1) // B
    res_list = [step_e2e(s_) for s_ in range(n_acc)][1:] + [pipe.drain()]     # results arrive one call late
    for s_, (Rb, tb, nb) in enumerate(res_list):
        for j, i in enumerate(batch_idx(s_)):
            if i not in acc_rows:
                acc_rows[i] = (Rb[j].numpy(), tb[j].numpy(), int(nb[j]))
    gpu_errs = [pose_errors(acc_rows[i][0], acc_rows[i][1], gt[i]) for i in sorted(acc_rows)]

    result = None
    if rank == 0:
        # ---- roofline of the dominant kernel (tcgen05 conv/GEMM), per-launch device times by CUDA events
        prof = eng.profile(dres["g0"][idx_cache[0]], dres["g1"][idx_cache[0]])
        tot_ms = sum(p[2] for p in prof)
        hbm, bf16, how, bf16_sus = load_peaks()
        traffic, traffic_src = load_traffic(args.precision, B)
        is_bb = lambda p: p[1] == 1 and p[0].startswith("layer")          # ResNet-FPN implicit-GEMM convs
        bb = [p for p in prof if is_bb(p)]
        og = [p for p in prof if p[1] == 1 and not is_bb(p)]             # transformer / correlation GEMMs (always TF32)
        bb_ms, bb_fl = sum(p[2] for p in bb), sum(p[3] for p in bb)
        og_ms, og_fl = sum(p[2] for p in og), sum(p[3] for p in og)
        kind = "bf16" if args.precision == "bf16" else "tf32"
        peak = bf16 if kind == "bf16" else bf16 / 2.0
        achieved = bb_fl / (bb_ms * 1e-3) / 1e12
        roofline = {"kernel": "conv_gemm_tf32_kernel, ResNet-FPN convs (tcgen05 kind::%s, %d launches per step)" % (
                        "f16/bf16" if kind == "bf16" else "tf32", len(bb)),
                    "bound": "tensor", "achieved": round(achieved, 1), "peak": round(peak, 1), "unit": "TFLOP/s",
                    "frac": round(achieved / peak, 4),
                    "peak_note": f"{how} cuBLAS bf16 {bf16} TFLOP/s" + ("" if kind == "bf16" else " / 2 (kind::tf32 issues at half the bf16 rate)"),
                    "frac_of_sustained_peak": round(achieved / (bf16_sus if kind == "bf16" else bf16_sus / 2.0), 4) if bf16_sus else None,
                    "share_of_step": round(bb_ms / tot_ms, 3), "flops_per_step": bb_fl, "traffic": traffic,
                    "traffic_note": (f"dram__bytes_read.sum + dram__bytes_write.sum summed over the {len(bb)} launches of one step, {traffic_src}"
                                     if traffic else "no ncu capture committed for this precision / batch"),
                    "other_gemms": {"launches": len(og), "achieved": round(og_fl / (og_ms * 1e-3) / 1e12, 1),
                                    "peak": round(bf16 / 2.0, 1), "share_of_step": round(og_ms / tot_ms, 3)}}
        cpu = cpu_baseline(args, sample=args.cpu_sample)
        fin = lambda v: [x for x in v if np.isfinite(x)]
        med = lambda v: round(float(np.median(v)), 4) if len(v) else None
        accuracy = {"against": "synthetic ground-truth poses of the pool pairs (EssentialMatrixMetric: R, metric t)",
                    "pairs": len(gpu_errs), "solved": len(fin([e[0] for e in gpu_errs])),
                    "median_rot_err_deg": med(fin([e[0] for e in gpu_errs])),
                    "median_trans_err_m": med(fin([e[1] for e in gpu_errs])),
                    "median_inliers": med([acc_rows[i][2] for i in sorted(acc_rows)]),
                    "vs_reference_cpu_port": (lambda d: {"pairs": len(d), "median_rot_diff_deg": med(fin([x[0] for x in d])),
                                                         "median_trans_diff_m": med(fin([x[1] for x in d])),
                                                         "note": "angle(R_gpu R_cpu^T), |t_gpu - t_cpu| on the same pairs; the two paths differ in matcher precision "
                                                                 "(bf16/tf32 vs fp32) and in the RANSAC (batched 5-pt + LO vs OpenCV USAC_MAGSAC)"})(
                        [pose_errors(acc_rows[i][0], acc_rows[i][1], CPU_POSES[i]) for i in range(min(len(CPU_POSES), n_pool)) if i in acc_rows]),
                    "reference_cpu_port": {"pairs": len(CPU_ERRS), "median_rot_err_deg": med(fin([e[0] for e in CPU_ERRS])),
                                           "median_trans_err_m": med(fin([e[1] for e in CPU_ERRS])),
                                           "median_matches": med([e[2] for e in CPU_ERRS]),
                                           "note": "same first pairs of the pool, oracle LoFTR fp32 + OpenCV USAC_MAGSAC"}}
        result = {
            "metric": "image-pairs/sec @ 540x720", "value": round(value, 2), "unit": "pairs/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": round(ms_res / K, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": ("bf16 backbone (bf16 storage, fp32 accumulate) + tf32 transformer/correlation; pose solve fp64/fp32"
                                         if args.precision == "bf16" else "tf32 (fp32 storage/accumulate; pose solve fp64/fp32)"), "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "pairs_per_step_per_gpu": B, "precision": args.precision, "engines": args.engines, "hypotheses_per_pair": args.hyp, "match_thr": args.match_thr,
                       "matches_last_batch": m_last, "weights": WEIGHTS,
                       "l2": "per-step working set (activations + 150 MB similarity matrix per pair) far exceeds the 126 MB L2; inputs rotate over %d distinct pairs" % n_pool},
            "e2e": {"value": round(e2e_value, 2), "unit": "pairs/s", "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": out_bytes, "ms_per_step": round(ms_e2e / K, 3),
                    "api": "mfr_b200.pipeline.RelocPipeline.submit_host (copy / matcher / RANSAC stages of consecutive batches overlap)"},
            "gpu_launches": int((eng.num_launches() + 5) * K),   # + emat_normalise, emat_hyp, emat_lo, emat_final, scale_consensus
            "kernels_per_step": eng.num_launches() + 5,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "accuracy": accuracy,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's algorithm restated in oracle/ (torch CPU LoFTR +
# OpenCV USAC_MAGSAC + scale consensus), all host threads
# ------------------------------------------------------------------------------------------------
CPU_ERRS = []   # (rot err deg, trans err m, matches) of every pair the CPU leg solved
CPU_POSES = []  # its (R, t) per pair, same order as the pool


def cpu_pairs(n_pairs, match_thr, seed0=1000):
    import torch
    from mfr_b200 import loftr
    from oracle import loftr_oracle as lo, pose_solver_oracle as po
    sd = loftr.functional_state_dict()
    g0, g1, d0, d1, Kc, gt = make_pool(n_pairs, seed0)
    t_net = t_solve = 0.0
    for i in range(n_pairs):
        i0 = loftr.pad_to_8(torch.from_numpy(g0[i])[None, None]); i1 = loftr.pad_to_8(torch.from_numpy(g1[i])[None, None])
        t0 = time.perf_counter()
        with torch.no_grad():
            o = lo.loftr_forward(i0, i1, sd, {"thr": match_thr})
        t1 = time.perf_counter()
        k0, k1 = o["mkpts0_f"].numpy(), o["mkpts1_f"].numpy()
        Rc, tc = po.essential_matrix_metric_solver(k0, k1, d0[i], d1[i], Kc[i], Kc[i], PIX_THR, 0.9999, SCALE_THR)[:2]
        t2 = time.perf_counter()
        t_net += t1 - t0; t_solve += t2 - t1
        CPU_ERRS.append(pose_errors(Rc, tc, gt[i]) + (len(k0),))
        CPU_POSES.append((np.asarray(Rc, np.float64), np.ravel(tc).astype(np.float64)))
    return t_net, t_solve


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use every core the process may run on."""
    import torch
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:                                   # one thread per physical core (SMT siblings only slow the GEMMs down)
        import psutil
        phys = psutil.cpu_count(logical=False)
        if phys:
            n = max(1, min(n, phys))
    except Exception:
        pass
    torch.set_num_threads(n)
    try:
        import cv2
        cv2.setNumThreads(n)
    except Exception:
        pass
    return torch.get_num_threads()


def cpu_baseline(args, sample):
    cores = use_all_host_threads()
    t_net, t_solve = cpu_pairs(sample, args.match_thr)
    return {"value": round(sample / (t_net + t_solve), 4), "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"{sample} pairs of the same workload: oracle LoFTR (torch CPU fp32, {cores} threads) {t_net / sample:.2f} s/pair + OpenCV USAC_MAGSAC 5-pt + scale consensus {t_solve / sample * 1e3:.1f} ms/pair"}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    K, Wm = args.steps, args.warmup
    cores = use_all_host_threads()
    if Wm > 0:
        cpu_pairs(1, args.match_thr)                           # warm-up (thread pools, allocator)
    t0 = time.perf_counter()
    t_net, t_solve = cpu_pairs(K, args.match_thr)              # one pair per step (bounded sample of a step)
    wall = time.perf_counter() - t0
    v = K / wall
    print(json.dumps({
        "impl": "reference", "metric": "image-pairs/sec @ 540x720", "value": round(v, 4), "unit": "pairs/s",
        "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": round(wall / K * 1e3, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (OpenCV solve f64)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "match_thr": args.match_thr, "weights": WEIGHTS,
                   "reference_sample": "reference CPU path (oracle port: torch CPU fp32 LoFTR + OpenCV USAC_MAGSAC + scale consensus), 1 pair per step on rank 0"},
        "cpu_baseline": {"value": round(v, 4), "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"{K} pairs, LoFTR {t_net / K:.2f} s/pair + solver {t_solve / K * 1e3:.1f} ms/pair"},
        "e2e": {"value": round(v, 4), "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "accuracy": {"against": "synthetic ground-truth poses", "pairs": len(CPU_ERRS),
                     "median_rot_err_deg": round(float(np.median([e[0] for e in CPU_ERRS])), 4),
                     "median_trans_err_m": round(float(np.median([e[1] for e in CPU_ERRS])), 4),
                     "median_matches": float(np.median([e[2] for e in CPU_ERRS]))}}))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
