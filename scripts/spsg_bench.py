"""BASELINE config C3 timing: SuperPoint + SuperGlue + PnP-from-depth, 540x720, B pairs per step (synthetic weights:
the reference's in-tree SuperPoint/SuperGlue weights do not travel to the GPU box; with random weights SuperPoint
fires on up to 1024 keypoints per image, so the SuperGlue workload is the full-size one)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mfr_b200
from mfr_b200 import spsg, synth, pose_solver

B = int(os.environ.get("B", 8)); H, W = 720, 540
sp, sg = spsg.synthetic_state_dicts(0)
eng = spsg.SuperPointSuperGlue(sp, sg, batch=B, height=H, width=W, graph=True)
be = pose_solver.PoseBackend(max_pairs=B, max_pts=1024, max_hyp=1024)
ps = [synth.make_pair(1000 + b) for b in range(B)]
g0 = torch.from_numpy(np.stack([synth.rgb_to_gray_u8(p["image0"]).astype(np.float32) / 255.0 for p in ps])).cuda()
g1 = torch.from_numpy(np.stack([synth.rgb_to_gray_u8(p["image1"]).astype(np.float32) / 255.0 for p in ps])).cuda()
d0 = torch.from_numpy(np.stack([p["depth0"] for p in ps]).astype(np.float32)).cuda()
K = torch.from_numpy(np.stack([p["K_color0"] for p in ps]).astype(np.float32)).cuda().contiguous()
outs = eng.device_outputs()

def step():
    eng.forward_async(g0, g1)
    with torch.cuda.stream(eng.stream):
        return be.pnp(outs["mkpts0"], outs["mkpts1"], outs["m_offsets"], K, K, d0, 3.0, 1000)

for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
with torch.cuda.stream(eng.stream):
    e0.record()
for _ in range(n):
    R, t, inl, mask = step()
with torch.cuda.stream(eng.stream):
    e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"SuperPoint+SuperGlue+PnP: B={B} pairs/step, {ms:.3f} ms/step, {B / ms * 1e3:.1f} pairs/s, kernels/step {eng.num_launches()}, "
      f"keypoints/image {outs['counts'].float().mean().item():.0f}, matches/pair {(outs['m_offsets'][1:] - outs['m_offsets'][:-1]).float().mean().item():.0f}")
