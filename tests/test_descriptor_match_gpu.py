"""Exact 2-NN descriptor search on the GPU vs cv2.BFMatcher (integer indices: bit-exact)."""
import cv2
import numpy as np
import pytest

from helpers import synth

pytestmark = pytest.mark.gpu


def _bf(des0, des1):
    m = cv2.BFMatcher(cv2.NORM_L2).knnMatch(des0, des1, k=2)
    i1 = np.array([a.trainIdx for a, b in m]); i2 = np.array([b.trainIdx for a, b in m])
    d1 = np.array([a.distance for a, b in m]); d2 = np.array([b.distance for a, b in m])
    return i1, d1, i2, d2


def test_knn2_random_descriptors():
    from mfr_b200 import model
    rng = np.random.default_rng(0)
    des0 = rng.random((700, 128), dtype=np.float32); des1 = rng.random((900, 128), dtype=np.float32)
    des1[17] = des1[400]                                   # an exact duplicate: lowest index must win
    i1, d1, i2, d2 = model.knn2_gpu(des0, des1)
    j1, e1, j2, e2 = _bf(des0, des1)
    assert np.array_equal(i1, j1) and np.array_equal(i2, j2)
    np.testing.assert_allclose(d1, e1, rtol=1e-5); np.testing.assert_allclose(d2, e2, rtol=1e-5)


def test_sift_ratio_matches_on_synthetic_pair():
    """C1-style plumbing: RootSIFT + exact 2-NN + ratio test gives the same match list as the CPU
    brute-force matcher, and the matches are geometrically consistent with the known warp."""
    from mfr_b200 import matchers, model
    p = synth.make_pair(3)
    g0 = synth.rgb_to_gray_u8(p["image0"]); g1 = synth.rgb_to_gray_u8(p["image1"])
    sift = cv2.SIFT_create(2048)
    kp0, des0 = sift.detectAndCompute(g0, None); kp1, des1 = sift.detectAndCompute(g1, None)
    r0, r1 = model.root_sift(des0), model.root_sift(des1)
    i1, d1, i2, d2 = model.knn2_gpu(r0, r1)
    j1, e1, j2, e2 = _bf(r0, r1)
    assert (i1 == j1).mean() > 0.999        # exact ties between RootSIFT rows may order differently
    good_gpu = d1 < 0.8 * d2; good_cpu = e1 < 0.8 * e2
    assert (good_gpu == good_cpu).mean() > 0.999
    pts = matchers.SIFT_matcher.match_arrays(sift, g0, g1)
    assert pts.shape[1] == 4 and len(pts) > 50
