"""Per-step device times of one LoFTR forward (CUDA events inside the engine), grouped."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mfr_b200
from mfr_b200 import loftr, synth
B = int(os.environ.get("B", 4)); H, W = 720, 544
# WEIGHTS: functional (matcher that matches, 99.8 % zero weights), bench (the benchmark's: functional + dense 1e-3 perturbation), dense (random)
_w = os.environ.get('WEIGHTS', 'functional')
_sd = loftr.functional_state_dict() if _w == 'functional' else (loftr.functional_state_dict(noise=1e-3) if _w == 'bench' else loftr.synthetic_state_dict(0))
eng = loftr.LoFTR(_sd, batch=B, height=H, width=W, cfg={"thr": float(os.environ.get("THR", 0.2))}, precision=os.environ.get("PRECISION", "tf32"))
ps = [synth.make_pair(1000 + b) for b in range(B)]
g0 = torch.from_numpy(np.stack([synth.rgb_to_gray_u8(p["image0"]).astype(np.float32) / 255.0 for p in ps])).cuda()
g1 = torch.from_numpy(np.stack([synth.rgb_to_gray_u8(p["image1"]).astype(np.float32) / 255.0 for p in ps])).cuda()
eng.forward(g0, g1)
prof = eng.profile(g0, g1)
prof = eng.profile(g0, g1)
tot = sum(x[2] for x in prof)
print(f"B={B}: total {tot:.3f} ms ({tot/B:.3f} ms/pair), steps {len(prof)}, matches {int(eng.device_outputs()['m_total'][0].item())}")
sec = collections.OrderedDict()
def section(i, name):
    if name.startswith("layer") or name in ("stem",): return "backbone"
    return name
groups = collections.OrderedDict()
for i, (name, kind, ms, fl) in enumerate(prof):
    key = name if kind == 1 else "simt"
    # coarse grouping by weight-name prefix
    if name.startswith("layer") and kind == 1: key = "backbone conv"
    elif name.startswith("coarse") or name == "qkv": key = "transformer gemm"
    elif name.startswith("fine"): key = "fine gemm"
    g = groups.setdefault(key, [0, 0.0, 0.0]); g[0] += 1; g[1] += ms; g[2] += fl
for k, (n, ms, fl) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:20s} x{n:3d} {ms:8.3f} ms  {fl/ms/1e9 if ms>0 else 0:8.1f} TF/s")
simt = collections.OrderedDict()
for name, kind, ms, fl in prof:
    if kind != 1:
        g = simt.setdefault(name, [0, 0.0]); g[0] += 1; g[1] += ms
print("--- SIMT steps by name")
for k, (n, ms) in sorted(simt.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:24s} x{n:3d} {ms*1e3:9.1f} us")
print("--- slowest steps")
for i in sorted(range(len(prof)), key=lambda i: -prof[i][2])[:int(os.environ.get('TOP', 28))]:
    n, k, ms, fl = prof[i]
    print(f"{i:4d} {n:24s} kind {k} {ms*1e3:9.1f} us {fl/ms/1e9 if ms>0 else 0:8.1f} TF/s")

# pose back end on the engine's matches
from mfr_b200 import pose_solver
be = pose_solver.PoseBackend(max_pairs=B, max_pts=(H // 8) * (W // 8), max_hyp=1024)
outs = eng.device_outputs()
d0 = torch.from_numpy(np.stack([p["depth0"] for p in ps]).astype(np.float32)).cuda(); d1 = torch.from_numpy(np.stack([p["depth1"] for p in ps]).astype(np.float32)).cuda()
K = torch.from_numpy(np.stack([p["K_color0"] for p in ps]).astype(np.float32)).cuda().contiguous()
for hyp in (256, 1024):
    with torch.cuda.stream(eng.stream):
        for _ in range(3):
            R, t, n = be.emat_metric(outs["mkpts0_f"], outs["mkpts1_f"], outs["m_offsets"], K, K, d0, d1, 2.0, 0.3, hyp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            R, t, n = be.emat_metric(outs["mkpts0_f"], outs["mkpts1_f"], outs["m_offsets"], K, K, d0, d1, 2.0, 0.3, hyp)
        e1.record()
    torch.cuda.synchronize()
    print(f"pose back end ({hyp} hyp, {B} pairs, matches/pair {outs['m_offsets'].diff().tolist()}): {e0.elapsed_time(e1)/10:.3f} ms, inliers {n.tolist()}")
