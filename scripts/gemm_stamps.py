"""In-kernel milestones + barrier wait times of the tcgen05 conv/GEMM kernel for a few LoFTR shapes
(block 0; MFR_CONV_FORCE="halo,MT,TW" selects the tiling)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mfr_b200
from mfr_b200 import lib as mfr
lib = mfr.load()
SHAPES = ((8, 360, 272, 128, 128, 3, 1), (8, 180, 136, 256, 256, 3, 1), (1, 1, 48960, 256, 256, 1, 1), (1, 1, 48960, 512, 512, 1, 1), (8, 360, 272, 128, 196, 3, 2), (8, 360, 272, 192, 196, 3, 1), (8, 360, 272, 196, 196, 3, 1), (8, 180, 136, 196, 196, 3, 1), (8, 360, 272, 196, 128, 3, 1))
for bf in [int(a) for a in os.environ.get('BF', '0,1').split(',')]:
    for res in [int(a) for a in os.environ.get('RES', '0,1').split(',')]:
        for (N, H, W, cin, cout, R, s) in SHAPES:
            dt = torch.bfloat16 if bf else torch.float32
            x = torch.randn(N, H, W, cin, device="cuda").to(dt); w = (torch.randn(cout, R * R, cin, device="cuda") * 0.05).to(dt)
            out = torch.empty(N, H // s, W // s, cout, device="cuda", dtype=dt)
            r = torch.randn(N, H // s, W // s, cout, device="cuda").to(dt) if res else None
            us = lib.mfr_debug_time_conv(mfr.ptr(x), N, H, W, cin, mfr.ptr(w), cout, R, s, mfr.ptr(out), 20, 3 if bf else 0, mfr.ptr(r))
            fl = 2.0 * N * (H // s) * (W // s) * cout * cin * R * R
            print(f"bf16={bf} res={res} {(N,H,W,cin,cout,R,s)}: {us:.1f} us/launch  {fl/us/1e6:.0f} TF/s", flush=True)
