"""The three-stream pipeline (mfr_b200.pipeline.RelocPipeline) must return, one call late, exactly the poses the
synchronous matcher -> solver sequence gives for the same batches."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipeline_equals_sequential():
    from mfr_b200 import loftr, pose_solver, synth
    from mfr_b200.pipeline import RelocPipeline
    B, H, W = 2, 240, 320
    K0 = np.array([[300.0, 0, 159.5], [0, 300.0, 119.5], [0, 0, 1]], np.float32)
    pairs = [synth.make_pair(40 + i, h=H, w=W, K=K0.astype(np.float64)) for i in range(6)]
    gray = lambda p, k: torch.from_numpy(synth.rgb_to_gray_u8(p[k]).astype(np.float32) / 255.0)
    batches = []
    for s in range(3):
        ps = pairs[2 * s:2 * s + 2]
        batches.append(tuple(torch.stack(x).pin_memory() for x in (
            [gray(p, "image0") for p in ps], [gray(p, "image1") for p in ps],
            [torch.from_numpy(p["depth0"]) for p in ps], [torch.from_numpy(p["depth1"]) for p in ps],
            [torch.from_numpy(p["K_color0"]) for p in ps])))
    eng = loftr.LoFTR(loftr.functional_state_dict(), batch=B, height=H, width=W, cfg={"thr": 0.2}, graph=True)
    be = pose_solver.PoseBackend(max_pairs=B, max_pts=(H // 8) * (W // 8), max_hyp=512)
    # sequential reference
    seq = []
    outs = eng.device_outputs()
    for g0, g1, d0, d1, K in batches:
        eng.forward_async(g0.cuda(), g1.cuda())
        with torch.cuda.stream(eng.stream):
            R, t, n = be.emat_metric(outs["mkpts0_f"], outs["mkpts1_f"], outs["m_offsets"], K.cuda(), K.cuda(),
                                     d0.cuda(), d1.cuda(), 2.0, 0.3, 512)
        torch.cuda.synchronize()
        seq.append((R.cpu(), t.cpu(), n.cpu()))
    assert all(int(n.min()) > 20 for _, _, n in seq), "the synthetic pairs should give solvable poses"
    pipe = RelocPipeline(eng, be, 2.0, 0.3, 512, depth_hw=(H, W))
    got = [pipe.submit_host(*b) for b in batches]
    assert got[0] is None
    res = got[1:] + [pipe.drain()]
    for (R, t, n), (Rs, ts, ns) in zip(res, seq):
        assert torch.equal(n, ns)
        assert torch.allclose(R, Rs, atol=1e-6) and torch.allclose(t, ts, atol=1e-6)
    eng.close(); be.close()
