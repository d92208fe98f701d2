"""Tiny driver for ncu captures of the conv/GEMM kernel: argv = N H W Cin Cout R stride iters [flags]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mfr_b200
from mfr_b200 import lib as mfr
lib = mfr.load()
N, H, W, Cin, Cout, R, s, iters = [int(a) for a in sys.argv[1:9]]
flags = int(sys.argv[9]) if len(sys.argv) > 9 else 0
dt = torch.bfloat16 if flags & 1 else torch.float32
x = torch.randn(N, H, W, Cin, device="cuda").to(dt); w = (torch.randn(Cout, R * R, Cin, device="cuda") * 0.05).to(dt)
Ho = (H + 2 * (R // 2) - R) // s + 1; Wo = (W + 2 * (R // 2) - R) // s + 1
out = torch.empty(N, Ho, Wo, Cout, device="cuda", dtype=torch.bfloat16 if flags & 2 else torch.float32)
us = lib.mfr_debug_time_conv(mfr.ptr(x), N, H, W, Cin, mfr.ptr(w), Cout, R, s, mfr.ptr(out), iters, flags, None)
print(f"{sys.argv[1:8]}: {us:.1f} us/launch, {2.0*N*Ho*Wo*Cout*Cin*R*R/us/1e6:.1f} TF/s")
