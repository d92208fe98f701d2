import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mfr_b200
from mfr_b200 import lib as mfr
lib = mfr.load()
for (N, H, W, cin, cout, R, s) in ((1, 1, 128, 256, 256, 1, 1), (1, 1, 6120, 256, 256, 1, 1), (1, 1, 6120, 512, 512, 1, 1), (2, 360, 272, 128, 128, 3, 1), (1, 1, 48960, 256, 256, 1, 1)):
    x = torch.randn(N, H, W, cin, device="cuda"); w = torch.randn(cout, R * R, cin, device="cuda") * 0.05
    out = torch.empty(N, H // s, W // s, cout, device="cuda")
    us = lib.mfr_debug_time_conv(mfr.ptr(x), N, H, W, cin, mfr.ptr(w), cout, R, s, mfr.ptr(out), 50)
    print(f"{(N,H,W,cin,cout,R,s)}: {us:.1f} us/launch", flush=True)
