"""Device-side leaderboard metrics (mfr_b200.metrics: CUDA kernel + torch aggregation) against the numpy oracle that
is pinned to the reference's benchmark/ code (tests/test_metrics_oracle.py)."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as mo

pytestmark = pytest.mark.gpu


def _poses(n, seed):
    rng = np.random.default_rng(seed)
    qg = np.stack([mo.euler2quat(*rng.uniform(0, 2 * np.pi, 3)) for _ in range(n)])
    qe = np.stack([mo.qmult(q, mo.axangle2quat(rng.uniform(-1, 1, 3), rng.uniform(0, 0.4) * (0.01 if i % 3 == 0 else 1.0))) for i, q in enumerate(qg)])
    qe[1] *= 3.0                                      # non-unit quaternion: both code paths normalise
    tg = rng.normal(0, 1, (n, 3)); te = tg + rng.normal(0, 0.3, (n, 3)) * rng.uniform(0, 1, (n, 1))
    K = np.tile(np.array([[590.0, 0, 270.0], [0, 590.0, 360.0], [0, 0, 1]]), (n, 1, 1)); K[:, 0, 0] += rng.uniform(-20, 20, n)
    return qg, tg, qe, te, K, rng.integers(0, 400, n).astype(np.float64)


def test_pose_metrics_match_oracle():
    from mfr_b200 import metrics
    qg, tg, qe, te, K, conf = _poses(301, 0)
    dev = "cuda"
    got = metrics.pose_metrics(*[torch.from_numpy(a).to(dev) for a in (qg, tg, qe, te, K)], 540, 720)
    ref = [mo.pose_metrics(qg[i], tg[i], qe[i], te[i], K[i], 540, 720) for i in range(len(qg))]
    for k in ("trans_err", "rot_err", "reproj_err"):
        np.testing.assert_allclose(got[k].cpu().numpy(), [r[k] for r in ref], rtol=1e-9, atol=1e-9)


def test_aggregate_and_auc_match_oracle():
    from mfr_b200 import metrics
    dev = "cuda"
    scenes_np, scenes_t = {}, {}
    for s in range(3):
        qg, tg, qe, te, K, conf = _poses(50 + 7 * s, 10 + s)
        m = metrics.pose_metrics(*[torch.from_numpy(a).to(dev) for a in (qg, tg, qe, te, K)], 540, 720)
        m["confidence"] = torch.from_numpy(conf).to(dev)
        scenes_t[f"s{s}"] = m
        scenes_np[f"s{s}"] = {k: v.cpu().numpy() for k, v in m.items()}
    got, ref = metrics.aggregate_results(scenes_t, 5), mo.aggregate_results(scenes_np, 5)
    assert set(got) == set(ref)
    for k in ref:
        assert got[k] == pytest.approx(ref[k], rel=1e-9, abs=1e-12), k
    c = torch.tensor([7.0, 9, 2, 7], device=dev); tp = torch.tensor([0.0, 1, 1, 1], device=dev)      # ties in the ranking
    p, r, ap = metrics.precision_recall(c, tp, 1)
    po, ro, apo = mo.precision_recall([7, 9, 2, 7], [0, 1, 1, 1], 1)
    assert np.allclose(p.cpu().numpy(), po) and np.allclose(r.cpu().numpy(), ro) and ap == pytest.approx(apo)
