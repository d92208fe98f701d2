"""The estimator shell end to end on the GPU: FeatureMatchingModel built from yaml files with the reference's keys
(config/mapfree.yaml + config/matching/mapfree/*_dptkitti.yaml) and a `Precomputed` correspondence npz in the
reference's on-disk format, `forward(data)` with a dataset-shaped batch (lib/datasets/mapfree.py:250-268 after the
DataLoader's collate), `submission.predict` + `save_submission` (submission.py:33-65). Poses are compared with the
committed outputs of the UNMODIFIED reference shell on the same inputs (tests/golden/pose_solver_reference.npz)."""
import io
import zipfile

import numpy as np
import pytest
import torch

from helpers import GOLDEN, POSE_CASES, pose_case, rot_angle
from mfr_b200 import config, matchers, model, submission

pytestmark = pytest.mark.gpu

DATASET_YAML = ("DATASET:\n  DATA_SOURCE: 'MapFree'\n  DATA_ROOT: 'data/mapfree/'\n  SCENES: None\n"
                "  ESTIMATED_DEPTH: None\n  AUGMENTATION_TYPE: None\n  HEIGHT: 720\n  WIDTH: 540\n"
                "  MIN_OVERLAP_SCORE: 0.2\n  MAX_OVERLAP_SCORE: 0.7\n  QUERY_FRAME_COUNT: 1\n")
METHOD_YAML = {
    "EssentialMatrixMetric": ("MODEL: 'FeatureMatching'\nFEATURE_MATCHING: 'Precomputed'\nPOSE_SOLVER: 'EssentialMatrixMetric'\n"
                              "DATASET:\n  ESTIMATED_DEPTH: 'dptkitti'\nMATCHES_FILE_PATH: '{scene_root}/correspondences_LoFTR.npz'\n"
                              "EMAT_RANSAC:\n  PIX_THRESHOLD: 2.0\n  SCALE_THRESHOLD: 0.1\n  CONFIDENCE: 0.9999\n"),
    "PNP": ("MODEL: 'FeatureMatching'\nFEATURE_MATCHING: 'Precomputed'\nPOSE_SOLVER: 'PNP'\n"
            "DATASET:\n  ESTIMATED_DEPTH: 'dptkitti'\nMATCHES_FILE_PATH: '{scene_root}/correspondences_SG.npz'\n"
            "PNP:\n  RANSAC_ITER: 1000\n  REPROJECTION_INLIER_THRESHOLD: 3\n  CONFIDENCE: 0.9999\n"),
    "EssentialMatrix": ("MODEL: 'FeatureMatching'\nFEATURE_MATCHING: 'Precomputed'\nPOSE_SOLVER: 'EssentialMatrix'\n"
                        "MATCHES_FILE_PATH: '{scene_root}/correspondences_LoFTR.npz'\n"
                        "EMAT_RANSAC:\n  PIX_THRESHOLD: 2.0\n  SCALE_THRESHOLD: 0.1\n  CONFIDENCE: 0.9999\n"),
}
CASES = [0, 1, 2, 3, 5]          # golden pose cases with enough correspondences (case 4 has 4 points: failure path)


def _scene(tmp_path, solver):
    """One synthetic 'scene': the cases' correspondences stacked into the npz the offline matcher driver writes
    (utils.py:59-69: NaN-padded float64 [Npairs, maxN, 4], row = pair_id), yaml files, and the batches a
    DataLoader(batch_size=1) would yield."""
    (tmp_path / "mapfree.yaml").write_text(DATASET_YAML)
    (tmp_path / "method.yaml").write_text(METHOD_YAML[solver])
    cs = [pose_case(ci) for ci in range(len(POSE_CASES))]
    corr = matchers.stack_pts([np.concatenate([c["kpts0"], c["kpts1"]], 1).astype(np.float64) for c in cs])
    root = tmp_path / "s00001"
    root.mkdir()
    for name in ("LoFTR", "SG"):
        np.savez_compressed(root / f"correspondences_{name}.npz", correspondences=corr)
    cfg = config.load_cfg(str(tmp_path / "mapfree.yaml"), str(tmp_path / "method.yaml"))
    batches = []
    for ci, c in enumerate(cs):
        batches.append({
            "image0": torch.zeros(1, 3, 720, 540), "image1": torch.zeros(1, 3, 720, 540),
            "depth0": torch.from_numpy(c["depth0"])[None], "depth1": torch.from_numpy(c["depth1"])[None],
            "K_color0": torch.from_numpy(c["K_color0"])[None], "K_color1": torch.from_numpy(c["K_color1"])[None],
            "scene_id": ["s00001"], "scene_root": [str(root)], "pair_id": torch.tensor([ci]),
            "pair_names": (["seq0/frame_00000.jpg"], [f"seq1/frame_{5 * ci:05d}.jpg"]), "dataset_name": ["Mapfree"]})
    return cfg, cs, batches


@pytest.mark.parametrize("solver", ["EssentialMatrixMetric", "PNP", "EssentialMatrix"])
def test_forward_matches_reference_shell(tmp_path, solver):
    cfg, cs, batches = _scene(tmp_path, solver)
    m = model.build_model(cfg)
    G = np.load(GOLDEN + "/pose_solver_reference.npz")
    for ci in CASES:
        data = batches[ci]
        R, t = m(data)
        assert R.shape == (1, 3, 3) and t.shape == (1, 1, 3) and R.dtype == torch.float32 and t.dtype == torch.float32
        assert isinstance(data["inliers"], int) and data["inliers"] > 0
        Rn, tn = R[0].numpy(), t[0, 0].numpy()
        Rg, tg = G[f"c{ci}_{solver}_R"], G[f"c{ci}_{solver}_t"]
        c = cs[ci]
        if solver == "PNP":        # raw tolerance of north_star (4e-3 on the 1 px-noise case, see test_geometry_gpu)
            tol = 4e-3 if POSE_CASES[ci][3] >= 1.0 else 1e-3
            assert rot_angle(Rn, Rg) < tol and np.linalg.norm(tn - tg) < tol
        elif solver == "EssentialMatrixMetric":   # no worse than the reference against ground truth (+1e-3)
            assert rot_angle(Rn, c["R_gt"]) <= rot_angle(Rg, c["R_gt"]) + 1e-3
            assert np.linalg.norm(tn - c["t_gt"]) <= np.linalg.norm(tg - c["t_gt"]) + 1e-3
        else:
            assert abs(np.linalg.norm(tn) - 1) < 1e-5
            assert rot_angle(Rn, c["R_gt"]) <= rot_angle(Rg, c["R_gt"]) + 1e-3
    # failure path: 4 correspondences -> NaN pose, inliers 0 (pose_solver.py:30-33), shapes kept
    data = batches[4]
    R, t = m(data)
    if solver != "PNP":
        assert torch.isnan(R).all() and torch.isnan(t).all() and data["inliers"] == 0
    assert R.shape == (1, 3, 3) and t.shape == (1, 1, 3)


def test_predict_and_submission_file(tmp_path):
    cfg, cs, batches = _scene(tmp_path, "EssentialMatrixMetric")
    m = model.build_model(cfg)
    results = submission.predict(batches, m)
    poses = results["s00001"]
    assert [p.image_name for p in poses] == [f"seq1/frame_{5 * ci:05d}.jpg" for ci in CASES]   # the NaN frame is skipped
    out = tmp_path / "submission.zip"
    submission.save_submission(results, out)
    with zipfile.ZipFile(out) as z:
        assert z.namelist() == ["pose_s00001.txt"]
        lines = z.read("pose_s00001.txt").decode().split("\n")
    assert len(lines) == len(CASES)
    for line, p, ci in zip(lines, poses, CASES):
        tok = line.split(" ")
        assert len(tok) == 9 and tok[0] == p.image_name and all(len(x.split(".")[1]) == 6 for x in tok[1:8])
        q = np.array([float(x) for x in tok[1:5]]); t = np.array([float(x) for x in tok[5:8]])
        assert abs(np.linalg.norm(q) - 1) < 1e-5 and q[0] >= 0
        w, x, y, z_ = q
        Rq = np.array([[1 - 2 * (y * y + z_ * z_), 2 * (x * y - z_ * w), 2 * (x * z_ + y * w)],
                       [2 * (x * y + z_ * w), 1 - 2 * (x * x + z_ * z_), 2 * (y * z_ - x * w)],
                       [2 * (x * z_ - y * w), 2 * (y * z_ + x * w), 1 - 2 * (x * x + y * y)]])
        c = cs[ci]
        assert rot_angle(Rq, c["R_gt"]) < 2e-2 and np.linalg.norm(t - c["t_gt"]) < 6e-2      # a usable pose made it to the file
        assert int(tok[8]) == p.inliers > 0
