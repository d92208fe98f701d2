"""Per-step device times of one LoFTR forward (CUDA events inside the engine), grouped."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mfr_b200
from mfr_b200 import loftr, synth
B = int(os.environ.get("B", 4)); H, W = 720, 544
eng = loftr.LoFTR(loftr.synthetic_state_dict(0), batch=B, height=H, width=W, cfg={"thr": float(os.environ.get("THR", 0.2))}, precision=os.environ.get("PRECISION", "tf32"))
p = synth.make_pair(0)
g0 = torch.from_numpy(synth.rgb_to_gray_u8(p["image0"]).astype(np.float32) / 255.0)[None].repeat(B, 1, 1).cuda()
g1 = torch.from_numpy(synth.rgb_to_gray_u8(p["image1"]).astype(np.float32) / 255.0)[None].repeat(B, 1, 1).cuda()
eng.forward(g0, g1)
prof = eng.profile(g0, g1)
prof = eng.profile(g0, g1)
tot = sum(x[2] for x in prof)
print(f"B={B}: total {tot:.3f} ms ({tot/B:.3f} ms/pair), steps {len(prof)}")
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
