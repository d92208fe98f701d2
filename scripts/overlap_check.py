"""How well does the solver stream hide behind the matcher? Times, at the bench shape: the LoFTR graph alone, the pose
back end alone, and the three-stream pipeline (device-resident inputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mfr_b200  # noqa: F401
from mfr_b200 import loftr, synth, pose_solver, pipeline
B = int(os.environ.get("B", 16)); H, W = 720, 544; N = 20
eng = loftr.LoFTR(loftr.functional_state_dict(), batch=B, height=H, width=W, cfg={"thr": 0.2}, precision=os.environ.get("PRECISION", "bf16"))
ps = [synth.make_pair(1000 + b) for b in range(B)]
g0 = torch.from_numpy(np.stack([synth.rgb_to_gray_u8(p["image0"]).astype(np.float32) / 255.0 for p in ps])).cuda()
g1 = torch.from_numpy(np.stack([synth.rgb_to_gray_u8(p["image1"]).astype(np.float32) / 255.0 for p in ps])).cuda()
d0 = torch.from_numpy(np.stack([p["depth0"] for p in ps]).astype(np.float32)).cuda()
d1 = torch.from_numpy(np.stack([p["depth1"] for p in ps]).astype(np.float32)).cuda()
K = torch.from_numpy(np.stack([p["K_color0"] for p in ps]).astype(np.float32)).cuda().contiguous()
be = pose_solver.PoseBackend(max_pairs=B, max_pts=eng.capacity // B, max_hyp=1024)
outs = eng.device_outputs()

def timed(fn, stream):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream): e0.record()
    for _ in range(N): fn()
    with torch.cuda.stream(stream): e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / N

fwd = timed(lambda: eng.forward_async(g0, g1), eng.stream)
def pose():
    with torch.cuda.stream(eng.stream):
        be.emat_metric(outs["mkpts0_f"], outs["mkpts1_f"], outs["m_offsets"], K, K, d0, d1, 2.0, 0.1, 1024)
pos = timed(pose, eng.stream)
pipe = pipeline.RelocPipeline(eng, be, 2.0, 0.1, 1024)
for _ in range(3): pipe.submit_device(g0, g1, d0, d1, K)
pipe.drain(); torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(N): pipe.submit_device(g0, g1, d0, d1, K)
pipe.drain(); torch.cuda.synchronize()
pp = (time.perf_counter() - t0) * 1e3 / N
print(f"B={B}: matcher graph alone {fwd:.2f} ms, pose back end alone {pos:.2f} ms, pipeline {pp:.2f} ms/step "
      f"(serial sum {fwd + pos:.2f}; overlap hides {100 * (fwd + pos - pp) / pos:.0f} % of the solver)")
