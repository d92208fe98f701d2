"""Stage-wise comparison of the CUDA LoFTR engine against the torch CPU oracle (prints errors)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mfr_b200
from mfr_b200 import loftr, synth
from oracle import loftr_oracle as lo

torch.manual_seed(0)
B = int(os.environ.get("B", 2)); H = int(os.environ.get("H", 240)); W = int(os.environ.get("W", 184))
thr = float(os.environ.get("THR", 0.0))
sd = lo.make_state_dict(0)
imgs0, imgs1 = [], []
for b in range(B):
    p = synth.make_pair(b)
    g0 = synth.rgb_to_gray_u8(p["image0"]).astype(np.float32) / 255.0
    g1 = synth.rgb_to_gray_u8(p["image1"]).astype(np.float32) / 255.0
    if H + 8 <= g0.shape[0] and W + 16 <= g0.shape[1] and b % 2 == 0:
        imgs0.append(g0[:H, :W]); imgs1.append(g0[8:H + 8, 16:W + 16])
    else:
        pad = lambda a: np.pad(a, ((0, max(0, H - a.shape[0])), (0, max(0, W - a.shape[1]))))[:H, :W]
        imgs0.append(pad(g0)); imgs1.append(pad(g1))
i0 = torch.from_numpy(np.stack(imgs0))[:, None]; i1 = torch.from_numpy(np.stack(imgs1))[:, None]
t0 = time.time()
with torch.no_grad():
    o = lo.loftr_forward(i0, i1, sd, {"thr": thr}, True)
print("oracle time %.2fs, M=%d" % (time.time() - t0, len(o["mkpts0_f"])))
eng = loftr.LoFTR(sd, batch=B, height=H, width=W, cfg={"thr": thr}, graph=bool(int(os.environ.get("GRAPH", 1))), precision=os.environ.get("PRECISION", "tf32"))
print("kernels per forward:", eng.num_kernels())
out = eng.forward(i0.cuda(), i1.cuda())
out = eng.forward(i0.cuda(), i1.cuda())  # second call exercises the graph replay
h8, w8, L = H // 8, W // 8, (H // 8) * (W // 8)
def rel(a, b):
    a = a.float().cpu(); b = b.float().cpu()
    return ((a - b).abs().max() / b.abs().max()).item(), b.abs().max().item()
with torch.no_grad():
    fcb, ffb = lo.backbone(torch.cat([i0, i1], 0), sd)
print("feat_c_backbone rel err", rel(eng.debug_tensor("feat_c_backbone", (2 * B, h8, w8, 256)), fcb.permute(0, 2, 3, 1)))
print("feat_f rel err", rel(eng.debug_tensor("feat_f", (2 * B, H // 2, W // 2, 128)), ffb.permute(0, 2, 3, 1)))
tok = eng.debug_tensor("tok", (2 * B, L, 512))[:, :, :256]
print("coarse tok0 rel err", rel(tok[:B], o["feat_c0"]), "tok1", rel(tok[B:], o["feat_c1"]))
sim = eng.debug_tensor("sim", (B, L, L))
print("sim rel err", rel(sim, o["sim"]), sim[0, :2, :4].tolist(), o["sim"][0, :2, :4].tolist(), sim.shape, o["sim"].shape)
M = out["M"]
print("M engine", M, "oracle", len(o["mkpts0_f"]), "offsets", out["m_offsets"].tolist())
e_set = set(zip(out["b_ids"].tolist(), out["i_ids"].tolist(), out["j_ids"].tolist()))
o_set = set(zip(o["b_ids"].tolist(), o["i_ids"].tolist(), o["j_ids"].tolist()))
inter = e_set & o_set
print("match-set IoU", len(inter) / max(1, len(e_set | o_set)), "common", len(inter))
# ordering check: engine output sorted by (b, i)
keys = [(b, i) for b, i in zip(out["b_ids"].tolist(), out["i_ids"].tolist())]
print("sorted by (b,i):", keys == sorted(keys))
# fine-level error on common matches
eo = {k: n for n, k in enumerate(zip(out["b_ids"].tolist(), out["i_ids"].tolist(), out["j_ids"].tolist()))}
oo = {k: n for n, k in enumerate(zip(o["b_ids"].tolist(), o["i_ids"].tolist(), o["j_ids"].tolist()))}
if inter:
    ie = torch.tensor([eo[k] for k in inter]); io = torch.tensor([oo[k] for k in inter])
    d1 = (out["mkpts1_f"].cpu()[ie] - o["mkpts1_f"][io]).abs()
    d0 = (out["mkpts0_f"].cpu()[ie] - o["mkpts0_f"][io]).abs()
    dc = (out["mconf"].cpu()[ie] - o["mconf"][io]).abs()
    print("mkpts0 max diff", d0.max().item(), "mkpts1_f max/median diff px", d1.max().item(), d1.median().item(),
          "mconf max diff", dc.max().item())
# timing
torch.cuda.synchronize()
a, b_ = i0.cuda(), i1.cuda()
for _ in range(3):
    eng.forward_async(a, b_)
eng.stream.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(eng.stream):
    e0.record()
    for _ in range(10):
        eng.forward_async(a, b_)
    e1.record()
eng.stream.synchronize()
print("ms per forward (B=%d, %dx%d): %.3f" % (B, H, W, e0.elapsed_time(e1) / 10))
