"""Stages reference-derived artefacts into oracle/_ref/ (git-ignored; travels to the GPU box with the gpurun
snapshot). TEST INFRASTRUCTURE: nothing under oracle/ is imported by the product path.

The reference is pure Python (no compilable sources for this path), so there is nothing to compile.
What the GPU-side parity tests need from the reference tree, and cannot get on the GPU box (no /root/reference
there), are DATA assets, copied verbatim:
  * the SuperPoint / SuperGlue weights shipped in-tree
    (etc/feature_matching_baselines/SuperGlue/models/weights/{superpoint_v1,superglue_indoor}.pth), which the
    SuperGlue README's known-answer test (README.md:121-127) and BASELINE config C3 are defined on;
  * the README's ScanNet sample pair (assets/scannet_sample_images/scene0711_00_frame-00{1680,1995}.jpg).
No reference SOURCE file is copied. Run in the build container: python oracle/build_ref.py
"""
import os
import shutil

REF = os.environ.get("MFR_REFERENCE", "/root/reference")
FMB = os.path.join(REF, "etc", "feature_matching_baselines")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

ASSETS = [
    ("SuperGlue/models/weights/superpoint_v1.pth", "weights/superpoint_v1.pth"),
    ("SuperGlue/models/weights/superglue_indoor.pth", "weights/superglue_indoor.pth"),
    ("SuperGlue/assets/scannet_sample_images/scene0711_00_frame-001680.jpg", "data/scene0711_00_frame-001680.jpg"),
    ("SuperGlue/assets/scannet_sample_images/scene0711_00_frame-001995.jpg", "data/scene0711_00_frame-001995.jpg"),
]


def available():
    return all(os.path.exists(os.path.join(OUT, dst)) for _, dst in ASSETS)


def weights_dir():
    return os.path.join(OUT, "weights")


def data_dir():
    return os.path.join(OUT, "data")


def build(verbose=True):
    if not os.path.isdir(FMB):
        if verbose:
            print("[oracle/build_ref] reference tree not mounted; keeping whatever oracle/_ref holds")
        return available()
    for src, dst in ASSETS:
        s, d = os.path.join(FMB, src), os.path.join(OUT, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.exists(d) or os.path.getsize(d) != os.path.getsize(s):
            shutil.copyfile(s, d)
            if verbose:
                print(f"[oracle/build_ref] {src} -> oracle/_ref/{dst}")
    return available()


if __name__ == "__main__":
    print("oracle/_ref complete:", build())
