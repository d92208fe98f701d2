"""ORACLE support (build container only): imports the UNMODIFIED reference modules from
/root/reference with the four shims of SURVEY.md §8(c). /root/reference does not exist on the GPU
box, so nothing under tests/ (gpu marker), smoke() or bench.py may import this file; it is used by
``tests/golden/make_golden.py`` and by the not-gpu tests that validate the restatements (skipped
when the reference tree is absent)."""
import os
import sys
import types

REF = os.environ.get("MFR_REFERENCE", "/root/reference")
FMB = os.path.join(REF, "etc", "feature_matching_baselines")


def available():
    return os.path.isdir(os.path.join(REF, "lib", "models", "matching"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_pose_solver():
    """reference lib.models.matching.{pose_solver,feature_matching,model}; shims: stub open3d,
    wrap backproject_3d so a torch K is np.asarray'd (numpy 2.x refuses ndarray * Tensor)."""
    _stub("open3d")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import numpy as np
    import lib.models.matching.pose_solver as ps
    if not getattr(ps, "_mfr_shimmed", False):
        orig = ps.backproject_3d

        def backproject_3d(uv, depth, K):
            out = orig(np.asarray(uv), np.asarray(depth), np.asarray(K))
            if hasattr(depth, "numpy"):  # PnPSolver calls .numpy() on the result (pose_solver.py:206)
                import torch
                return torch.from_numpy(out)
            return out

        ps.backproject_3d = backproject_3d
        ps._mfr_shimmed = True
    return ps


def load_model_module():
    load_pose_solver()
    import lib.models.matching.model as model
    return model


class Cfg(dict):
    """Minimal attribute-dict standing in for the yacs CfgNode (yacs is absent here)."""
    __getattr__ = dict.__getitem__

    @staticmethod
    def wrap(d):
        return Cfg({k: Cfg.wrap(v) if isinstance(v, dict) else v for k, v in d.items()})


def _yacs_stub():
    class CN(dict):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

    cfgmod = _stub("yacs.config", CfgNode=CN)
    _stub("yacs", config=cfgmod)


def _kornia_stub():
    import torch

    def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
        xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
        ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
        if normalized_coordinates:
            xs = (xs / (width - 1) - 0.5) * 2
            ys = (ys / (height - 1) - 0.5) * 2
        base = torch.stack(torch.meshgrid([xs, ys], indexing="ij"), dim=-1)
        return base.permute(1, 0, 2).unsqueeze(0)

    def spatial_expectation2d(inp, normalized_coordinates=True):
        b, c, h, w = inp.shape
        grid = create_meshgrid(h, w, normalized_coordinates, inp.device).to(inp.dtype)
        pos_x = grid[..., 0].reshape(-1)
        pos_y = grid[..., 1].reshape(-1)
        flat = inp.view(b, c, -1)
        ex = torch.sum(pos_x * flat, -1, keepdim=True)
        ey = torch.sum(pos_y * flat, -1, keepdim=True)
        return torch.cat([ex, ey], -1).view(b, c, 2)

    dsnt = _stub("kornia.geometry.subpix.dsnt", spatial_expectation2d=spatial_expectation2d)
    subpix = _stub("kornia.geometry.subpix", dsnt=dsnt)
    geometry = _stub("kornia.geometry", subpix=subpix)
    grid = _stub("kornia.utils.grid", create_meshgrid=create_meshgrid)
    utils = _stub("kornia.utils", grid=grid)
    _stub("kornia", geometry=geometry, utils=utils)


def load_loftr():
    """reference LoFTR (inference subset); shims: yacs.config.CfgNode, two kornia helpers."""
    _yacs_stub()
    _kornia_stub()
    if FMB not in sys.path:
        sys.path.insert(0, FMB)
    from LoFTR.src.loftr import LoFTR, default_cfg
    return LoFTR, default_cfg


def load_superglue():
    """reference SuperPoint / SuperGlue / Matching with the in-tree weights."""
    if FMB not in sys.path:
        sys.path.insert(0, FMB)
    from SuperGlue.models.superpoint import SuperPoint
    from SuperGlue.models.superglue import SuperGlue
    from SuperGlue.models.matching import Matching
    return SuperPoint, SuperGlue, Matching
