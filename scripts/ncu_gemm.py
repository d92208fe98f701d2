"""Tiny driver for ncu captures of the conv/GEMM kernel: argv = N H W Cin Cout R stride iters"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mfr_b200
from mfr_b200 import lib as mfr
lib = mfr.load()
N, H, W, Cin, Cout, R, s, iters = [int(a) for a in sys.argv[1:9]]
x = torch.randn(N, H, W, Cin, device="cuda"); w = torch.randn(Cout, R * R, Cin, device="cuda") * 0.05
Ho = (H + 2 * (R // 2) - R) // s + 1; Wo = (W + 2 * (R // 2) - R) // s + 1
out = torch.empty(N, Ho, Wo, Cout, device="cuda")
us = lib.mfr_debug_time_conv(mfr.ptr(x), N, H, W, Cin, mfr.ptr(w), Cout, R, s, mfr.ptr(out), iters)
print(f"{sys.argv[1:8]}: {us:.1f} us/launch, {2.0*N*Ho*Wo*Cout*Cin*R*R/us/1e6:.1f} TF/s")
