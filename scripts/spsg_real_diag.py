"""Diagnostics: SuperPoint/SuperGlue engine with the real weights vs the reference golden, per precision."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import GOLDEN, spsg_real_cases
from oracle import build_ref
from mfr_b200 import spsg
w = build_ref.weights_dir()
sp, sg = torch.load(w + "/superpoint_v1.pth", map_location="cpu"), torch.load(w + "/superglue_indoor.pth", map_location="cpu")
G = np.load(GOLDEN + "/spsg_real_reference.npz")
for name, i0, i1 in spsg_real_cases():
    for prec in ("fp32x3", "tf32"):
        H, W = i0.shape[-2:]
        eng = spsg.SuperPointSuperGlue(sp, sg, batch=1, height=H, width=W, graph=False, precision=prec)
        o = eng.forward(i0.cuda(), i1.cuda())
        for side in (0, 1):
            n = int(o["counts"][side])
            k = o["keypoints"][side, :n].cpu().numpy().astype(int); s = o["scores"][side, :n].cpu().numpy()
            gk = G[f"{name}_keypoints{side}"].astype(int); gs = G[f"{name}_scores{side}"]
            got = {tuple(x): float(v) for x, v in zip(k.tolist(), s.tolist())}; exp = {tuple(x): float(v) for x, v in zip(gk.tolist(), gs.tolist())}
            only_g = sorted(set(got) - set(exp)); only_e = sorted(set(exp) - set(got))
            common = sorted(set(got) & set(exp))
            ds = max(abs(got[c] - exp[c]) for c in common) if common else -1
            print(name, prec, "img", side, "n", n, "ref", len(gk), "only_gpu", [(c, round(got[c], 5)) for c in only_g][:6], "only_ref", [(c, round(exp[c], 5)) for c in only_e][:6], "max score diff", ds, flush=True)
        n0 = int(o["counts"][0])
        m0 = o["matches0"][0, :n0].cpu().numpy()
        k0 = o["keypoints"][0, :n0].cpu().numpy().astype(int); k1 = o["keypoints"][1, :int(o["counts"][1])].cpu().numpy().astype(int)
        gm = {(tuple(k0[i]), tuple(k1[j])) for i, j in enumerate(m0) if j > -1}
        em = {(tuple(G[name + "_keypoints0"][i].astype(int)), tuple(G[name + "_keypoints1"][j].astype(int))) for i, j in enumerate(G[name + "_matches0"]) if j > -1}
        print(name, prec, "matches gpu", len(gm), "ref", len(em), "common", len(gm & em), flush=True)
        eng.close()
