"""The hand-constructed LoFTR weights used by bench.py (loftr.functional_state_dict) must make the
network a working matcher: checked on the CPU oracle here and on the CUDA engine (gpu marker)."""
import numpy as np
import pytest
import torch

from oracle import loftr_oracle as lo


def _shifted_pair(H=160, W=200, dy=8, dx=16):
    from mfr_b200 import synth
    p = synth.make_pair(3)
    g = synth.rgb_to_gray_u8(p["image0"]).astype(np.float32) / 255.0
    a = g[100:100 + H, 60:60 + W]
    b = g[100 + dy:100 + dy + H, 60 + dx:60 + dx + W]          # b(y, x) = a(y + dy, x + dx)
    return torch.from_numpy(np.ascontiguousarray(a))[None, None], torch.from_numpy(np.ascontiguousarray(b))[None, None]


def _check_matches(k0, k1, dy, dx, min_matches):
    assert len(k0) >= min_matches
    d = k1 - k0                                               # a point at x in image 0 sits at x - shift in image 1
    ok = (np.abs(d[:, 0] + dx) < 1.5) & (np.abs(d[:, 1] + dy) < 1.5)
    assert ok.mean() > 0.8, ok.mean()


def test_functional_weights_match_a_shifted_crop_oracle():
    from mfr_b200 import loftr
    sd = loftr.functional_state_dict()
    i0, i1 = _shifted_pair()
    with torch.no_grad():
        o = lo.loftr_forward(i0, i1, sd, {"thr": 0.2})
    _check_matches(o["mkpts0_f"].numpy(), o["mkpts1_f"].numpy(), 8, 16, 100)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_functional_weights_engine_matches_oracle(precision):
    from mfr_b200 import loftr
    sd = loftr.functional_state_dict()
    i0, i1 = _shifted_pair()
    with torch.no_grad():
        o = lo.loftr_forward(i0, i1, sd, {"thr": 0.2})
    eng = loftr.LoFTR(sd, batch=1, height=i0.shape[2], width=i0.shape[3], cfg={"thr": 0.2}, graph=False, precision=precision)
    out = eng.forward(i0.cuda(), i1.cuda())
    k0, k1 = out["mkpts0_f"].cpu().numpy(), out["mkpts1_f"].cpu().numpy()
    _check_matches(k0, k1, 8, 16, 100)
    e = set(zip(out["i_ids"].tolist(), out["j_ids"].tolist()))
    r = set(zip(o["i_ids"].tolist(), o["j_ids"].tolist()))
    assert len(e & r) >= 0.9 * len(e | r)                     # same coarse matches as the fp32 oracle
    eng.close()
