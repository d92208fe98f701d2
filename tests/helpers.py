import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

import mfr_b200  # noqa: E402,F401
from mfr_b200 import synth  # noqa: E402

POSE_CASES = [(11, 400, 0.6, 0.3), (12, 800, 0.5, 0.5), (13, 300, 0.8, 0.0), (14, 1000, 0.4, 1.0),
              (15, 4, 1.0, 0.0), (16, 600, 0.7, 0.2)]
GOLD_H, GOLD_W = 720, 540
GOLD_K = np.array([[600.0, 0, 269.5], [0, 600.0, 359.5], [0, 0, 1]])


def pose_case(ci):
    seed, n, ratio, noise = POSE_CASES[ci]
    return synth.make_correspondences(seed, n=n, inlier_ratio=ratio, noise_px=noise, h=GOLD_H, w=GOLD_W, K=GOLD_K)


def checksum(*arrays):
    return float(sum(np.float64(np.asarray(a, dtype=np.float64)).sum() for a in arrays))


def loftr_case(ci):
    import torch
    pair, thr, same = [(0, 0.0, True), (1, 0.0, False)][ci]
    p = synth.make_pair(pair)
    g0 = synth.rgb_to_gray_u8(p["image0"]).astype(np.float32) / 255.0
    g1 = synth.rgb_to_gray_u8(p["image1"]).astype(np.float32) / 255.0
    H, W = 128, 96
    i0 = torch.from_numpy(g0[100:100 + H, 100:100 + W].copy())[None, None]
    i1 = torch.from_numpy((g0[108:108 + H, 116:116 + W] if same else g1[100:100 + H, 100:100 + W]).copy())[None, None]
    return i0, i1, thr


def rot_angle(Ra, Rb):
    c = (np.trace(np.asarray(Ra, dtype=np.float64) @ np.asarray(Rb, dtype=np.float64).T) - 1.0) / 2.0
    return float(np.arccos(np.clip(c, -1.0, 1.0)))


def vec_angle(a, b):
    a = np.asarray(a, dtype=np.float64).ravel(); b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.arccos(np.clip(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)), -1.0, 1.0)))
