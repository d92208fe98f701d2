"""L2 (LTS) pressure check of the tcgen05 conv kernel: the 3x3 layers at the bench batch (32 images), default tiling
vs CTA-pair weight multicast (MFR_CONV_FORCE=-1,-1,-1,-1,1). One process per setting (the tiling is chosen at build)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    import mfr_b200  # noqa: F401
    from mfr_b200 import lib as mfr
    lib = mfr.load()
    SHAPES = ((32, 360, 272, 192, 192, 3, 1), (32, 180, 136, 256, 256, 3, 1), (32, 360, 272, 128, 128, 3, 1), (32, 90, 68, 256, 256, 3, 1))
    for (N, H, W, cin, cout, R, s) in SHAPES:
        x = torch.randn(N, H, W, cin, device="cuda").to(torch.bfloat16)
        w = (torch.randn(cout, R * R, cin, device="cuda") * 0.05).to(torch.bfloat16)
        out = torch.empty(N, H // s, W // s, cout, device="cuda", dtype=torch.bfloat16)
        us = lib.mfr_debug_time_conv(mfr.ptr(x), N, H, W, cin, mfr.ptr(w), cout, R, s, mfr.ptr(out), 20, 3, None)
        fl = 2.0 * N * (H // s) * (W // s) * cout * cin * R * R
        print(f"{(N, H, W, cin, cout, R, s)}: {us:.1f} us/launch  {fl / us / 1e6:.0f} TF/s", flush=True)
else:
    for force in ("-1,-1,-1,0,0", "-1,-1,-1,0,1"):
        print("MFR_CONV_FORCE=" + force, flush=True)
        env = dict(os.environ, MFR_CONV_FORCE=force, MFR_CONV_DEBUG="1")
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, check=False)
