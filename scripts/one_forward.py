"""ncu driver: ONE LoFTR forward + pose solve at bench shape after one untimed warm-up forward.
Prints the number of kernel launches of the warm-up so that ncu can --launch-skip it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mfr_b200
from mfr_b200 import loftr, synth, pose_solver
B = int(os.environ.get("B", 4)); H, W = 720, 544
# WEIGHTS=dense (default): the throughput benchmark's weights, functional matcher + dense perturbation (bench.py); functional: 99.8 % zeros
noise = 0.0 if os.environ.get("WEIGHTS", "dense") == "functional" else 1e-3
eng = loftr.LoFTR(loftr.functional_state_dict(noise=noise), batch=B, height=H, width=W, cfg={"thr": 0.2}, graph=False,
                  precision=os.environ.get("PRECISION", "bf16"))
be = pose_solver.PoseBackend(max_pairs=B, max_pts=(H // 8) * (W // 8), max_hyp=1024)
ps = [synth.make_pair(1000 + b) for b in range(B)]
g0 = torch.from_numpy(np.stack([synth.rgb_to_gray_u8(p["image0"]).astype(np.float32) / 255.0 for p in ps])).cuda()
g1 = torch.from_numpy(np.stack([synth.rgb_to_gray_u8(p["image1"]).astype(np.float32) / 255.0 for p in ps])).cuda()
d0 = torch.from_numpy(np.stack([p["depth0"] for p in ps]).astype(np.float32)).cuda()
d1 = torch.from_numpy(np.stack([p["depth1"] for p in ps]).astype(np.float32)).cuda()
K = torch.from_numpy(np.stack([p["K_color0"] for p in ps]).astype(np.float32)).cuda().contiguous()
outs = eng.device_outputs()
for it in range(2):
    eng.forward_async(g0, g1)
    with torch.cuda.stream(eng.stream):
        R, t, n = be.emat_metric(outs["mkpts0_f"], outs["mkpts1_f"], outs["m_offsets"], K, K, d0, d1, 2.0, 0.3, 1024)
    torch.cuda.synchronize()
print("launches per forward:", eng.num_launches())
