"""Submission writer with the reference's format (submission.py:18-65): one text file per scene,
``pose_{scene}.txt`` with lines ``<query image> qw qx qy qz tx ty tz <inliers>`` (6 decimals),
zipped flat. ``mat2quat`` restates transforms3d.quaternions.mat2quat (w, x, y, z; largest
eigenvector of the symmetric K matrix; w >= 0), the reference's only use of transforms3d here."""
from collections import defaultdict
from dataclasses import dataclass
from zipfile import ZipFile

import numpy as np
import torch


def mat2quat(M):
    Qxx, Qyx, Qzx, Qxy, Qyy, Qzy, Qxz, Qyz, Qzz = np.asarray(M, dtype=np.float64).flat
    K = np.array([
        [Qxx - Qyy - Qzz, 0, 0, 0],
        [Qyx + Qxy, Qyy - Qxx - Qzz, 0, 0],
        [Qzx + Qxz, Qzy + Qyz, Qzz - Qxx - Qyy, 0],
        [Qyz - Qzy, Qzx - Qxz, Qxy - Qyx, Qxx + Qyy + Qzz]]) / 3.0
    vals, vecs = np.linalg.eigh(K)
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    if q[0] < 0:
        q *= -1
    return q


@dataclass
class Pose:
    image_name: str
    q: np.ndarray
    t: np.ndarray
    inliers: float

    def __str__(self) -> str:
        formatter = {"float": lambda v: f"{v:.6f}"}
        q_str = np.array2string(self.q, formatter=formatter, max_line_width=1000)[1:-1]
        t_str = np.array2string(self.t, formatter=formatter, max_line_width=1000)[1:-1]
        return f"{self.image_name} {q_str} {t_str} {self.inliers}"


def predict(loader, model):
    """submission.py:33-58: frames whose pose contains NaN/Inf are skipped."""
    results = defaultdict(list)
    for data in loader:
        with torch.no_grad():
            R, t = model(data)
        R = R.detach().cpu().numpy()
        t = t.reshape(-1).detach().cpu().numpy()
        if np.isnan(R).any() or np.isnan(t).any() or np.isinf(t).any():
            continue
        results[data["scene_id"][0]].append(
            Pose(image_name=data["pair_names"][1][0], q=mat2quat(R.reshape(3, 3)).reshape(-1), t=t.reshape(-1),
                 inliers=data["inliers"]))
    return results


def save_submission(results, output_path):
    with ZipFile(output_path, "w") as z:
        for scene, poses in results.items():
            z.writestr(f"pose_{scene}.txt", "\n".join(str(p) for p in poses).encode("utf-8"))
