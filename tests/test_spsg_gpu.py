"""GPU parity of the SuperPoint + SuperGlue engine against the torch fp32 CPU oracle
(oracle/spsg_oracle.py, itself pinned to the reference modules with the real in-tree weights).
Synthetic seeded weights (the real ones do not travel to the GPU box). TF32 tensor-core inputs:
dense maps within 2e-2 of max magnitude, keypoint sets IoU >= 0.9, SuperGlue on identical inputs:
score matrix within 2e-2, match sets IoU >= 0.9."""
import numpy as np
import pytest
import torch

from helpers import synth
from oracle import spsg_oracle as so

pytestmark = pytest.mark.gpu
H, W, K = 240, 320, 256


@pytest.fixture(scope="module")
def weights():
    from mfr_b200 import spsg
    return spsg.synthetic_state_dicts(0)


def _pair():
    p = synth.make_pair(5)
    g0 = synth.rgb_to_gray_u8(p["image0"]).astype(np.float32) / 255.0
    return torch.from_numpy(g0[40:40 + H, 30:30 + W].copy())[None, None], torch.from_numpy(g0[48:48 + H, 46:46 + W].copy())[None, None]


def _rel(a, b):
    a = a.float().cpu(); b = b.float().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


def test_superpoint_matches_oracle(weights):
    from mfr_b200 import spsg
    sp, sg = weights
    i0, i1 = _pair()
    cfg = {"max_keypoints": K}
    eng = spsg.SuperPointSuperGlue(sp, sg, batch=1, height=H, width=W, cfg=cfg, graph=False)
    assert eng.num_launches() > 150
    out = eng.forward(i0.cuda(), i1.cuda())
    h8, w8 = H // 8, W // 8
    for n, img in enumerate((i0, i1)):
        with torch.no_grad():
            sc, desc = so.superpoint_dense(img, sp)
            k, s, d = so.superpoint(img, sp, cfg)
        assert _rel(eng.debug_tensor("scores", (2, h8 * 8, w8 * 8))[n], sc[0]) < 2e-2
        assert _rel(eng.debug_tensor("dense_desc", (2, h8, w8, 256))[n], desc[0].permute(1, 2, 0)) < 2e-2
        cnt = int(out["counts"][n])
        got = {tuple(x) for x in out["keypoints"][n, :cnt].cpu().numpy().astype(int).tolist()}
        exp = {tuple(x) for x in k.numpy().astype(int).tolist()}
        assert len(got & exp) >= 0.9 * len(got | exp)
        # descriptors of common keypoints point the same way
        gi = {tuple(x): j for j, x in enumerate(out["keypoints"][n, :cnt].cpu().numpy().astype(int).tolist())}
        ei = {tuple(x): j for j, x in enumerate(k.numpy().astype(int).tolist())}
        common = sorted(got & exp)
        dg = out["descriptors"][n].cpu()[[gi[c] for c in common]]
        de = d.t()[[ei[c] for c in common]]
        assert (dg * de).sum(1).min().item() > 0.999
    eng.close()


def test_superglue_stage_matches_oracle(weights):
    """SuperGlue alone on identical (oracle) keypoints/descriptors, match threshold 0 so that every
    mutual nearest neighbour of the Sinkhorn assignment becomes a match."""
    from mfr_b200 import spsg
    sp, sg = weights
    i0, i1 = _pair()
    cfg = {"max_keypoints": K, "match_threshold": 0.0}
    with torch.no_grad():
        k0, s0, d0 = so.superpoint(i0, sp, cfg)
        k1, s1, d1 = so.superpoint(i1, sp, cfg)
        k1, s1, d1 = k1[:K - 37], s1[:K - 37], d1[:, :K - 37]          # ragged: n0 != n1
        m0, ms0, scores, Z = so.superglue(k0, s0, d0, k1, s1, d1, H, W, sg, cfg, return_scores=True)
    eng = spsg.SuperPointSuperGlue(sp, sg, batch=1, height=H, width=W, cfg=cfg, graph=False)
    counts = np.array([len(k0), len(k1)], np.int32)
    kp = np.zeros((2, K, 2), np.float32); sc = np.zeros((2, K), np.float32); de = np.zeros((2, K, 256), np.float32)
    kp[0, :len(k0)] = k0.numpy(); kp[1, :len(k1)] = k1.numpy()
    sc[0, :len(k0)] = s0.numpy(); sc[1, :len(k1)] = s1.numpy()
    de[0, :len(k0)] = d0.t().numpy(); de[1, :len(k1)] = d1.t().numpy()
    out = eng.forward_superglue(counts, kp, sc, de)
    S = eng.debug_tensor("sg_scores", (1, K, K))[0, :len(k0), :len(k1)]
    assert _rel(S, scores) < 2e-2
    got = out["matches0"][0, :len(k0)].cpu().long()
    assert (got >= -1).all() and (got < len(k1)).all()
    both = (got > -1) & (m0 > -1)
    union = ((got > -1) | (m0 > -1)).sum().item()
    agree = (both & (got == m0)).sum().item()
    assert union > 30 and agree >= 0.9 * union
    ms = out["matching_scores0"][0, :len(k0)].cpu()
    sel = both & (got == m0)
    assert (ms[sel] - ms0[sel]).abs().max().item() < 0.05
    # compacted coordinates follow matches0 order
    m = int(out["m_offsets"][1])
    idx = torch.nonzero(got > -1)[:, 0]
    assert m == len(idx)
    assert torch.equal(out["mkpts0"][:m].cpu(), k0[idx]) and torch.equal(out["mkpts1"][:m].cpu(), k1[got[idx]])
    eng.close()


def test_end_to_end_plugin_and_graph(weights):
    from mfr_b200 import spsg
    sp, sg = weights
    i0, i1 = _pair()
    m = spsg.SuperGlue_matcher.__new__(spsg.SuperGlue_matcher)
    m.resize = (W, H)
    m.engine = spsg.SuperPointSuperGlue(sp, sg, batch=1, height=H, width=W, cfg={"max_keypoints": K, "match_threshold": 0.0})
    a = m.match_arrays(i0[0, 0].numpy(), i1[0, 0].numpy())
    b = m.match_arrays(i0[0, 0].numpy(), i1[0, 0].numpy())     # second call = CUDA-graph replay
    assert a.shape[1] == 4 and len(a) > 10 and np.array_equal(a, b)
    m.engine.close()
