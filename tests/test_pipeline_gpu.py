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


def test_pipeline_separate_intrinsics_and_reset():
    """K0 / K1 travel separately (Map-free frames have per-frame intrinsics, lib/datasets/mapfree.py:39-54): the pipeline
    must give what the synchronous solver gives with the same (K0, K1); reset() starts a new stream."""
    from mfr_b200 import loftr, pose_solver, synth
    from mfr_b200.pipeline import RelocPipeline
    B, H, W = 2, 240, 320
    K0 = np.array([[300.0, 0, 159.5], [0, 300.0, 119.5], [0, 0, 1]], np.float32)
    pairs = [synth.make_pair(50 + i, h=H, w=W, K=K0.astype(np.float64)) for i in range(B)]
    gray = lambda p, k: torch.from_numpy(synth.rgb_to_gray_u8(p[k]).astype(np.float32) / 255.0)
    g0, g1 = torch.stack([gray(p, "image0") for p in pairs]), torch.stack([gray(p, "image1") for p in pairs])
    d0, d1 = torch.stack([torch.from_numpy(p["depth0"]) for p in pairs]), torch.stack([torch.from_numpy(p["depth1"]) for p in pairs])
    Ka = torch.from_numpy(np.stack([K0] * B))
    Kb = Ka.clone(); Kb[:, 0, 0] *= 1.04; Kb[:, 1, 1] *= 1.04; Kb[:, 0, 2] += 3.0      # a different camera for the query frames
    eng = loftr.LoFTR(loftr.functional_state_dict(), batch=B, height=H, width=W, cfg={"thr": 0.2}, graph=True)
    be = pose_solver.PoseBackend(max_pairs=B, max_pts=(H // 8) * (W // 8), max_hyp=512)
    outs = eng.device_outputs()
    eng.forward_async(g0.cuda(), g1.cuda())
    with torch.cuda.stream(eng.stream):
        Rs, ts, ns = be.emat_metric(outs["mkpts0_f"], outs["mkpts1_f"], outs["m_offsets"], Ka.cuda(), Kb.cuda(), d0.cuda(), d1.cuda(), 2.0, 0.3, 512)
    torch.cuda.synchronize()
    pipe = RelocPipeline(eng, be, 2.0, 0.3, 512, depth_hw=(H, W))
    batch = [x.pin_memory() for x in (g0, g1, d0, d1, Ka, Kb)]
    assert pipe.submit_host(*batch) is None
    R, t, n = pipe.drain()
    assert torch.equal(n, ns.cpu()) and torch.allclose(R, Rs.cpu(), atol=1e-6) and torch.allclose(t, ts.cpu(), atol=1e-6)
    pipe.reset()
    assert pipe.submit_host(*batch[:5]) is None                      # K1 omitted: both frames share K0 -> a different pose
    R2, t2, n2 = pipe.drain()
    assert not torch.allclose(t2, t, atol=1e-4)
    eng.close(); be.close()
