"""Fixed cost of one tcgen05 GEMM launch: probes + prebuilt ops launched back to back (CUDA events in C)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mfr_b200
from mfr_b200 import lib as mfr
lib = mfr.load()
torch.zeros(1, device="cuda")
for mode in (0, 1, 2):
    for smem in (0, 64 * 1024, 192 * 1024):
        for grid in (48, 148):
            print(f"probe mode {mode} smem {smem//1024}KB grid {grid}: {lib.mfr_debug_probe_overhead(mode, smem, grid, 200):.2f} us")
for rows in (128, 6120, 12240, 6120 * 8):
    for (cin, cout) in ((256, 256), (512, 512), (256, 768)):
        x = torch.randn(1, 1, rows, cin, device="cuda"); w = torch.randn(cout, 1, cin, device="cuda") * 0.05
        out = torch.empty(1, 1, rows, cout, device="cuda")
        us = lib.mfr_debug_time_conv(mfr.ptr(x), 1, 1, rows, cin, mfr.ptr(w), cout, 1, 1, mfr.ptr(out), 100, 0, None)
        print(f"prebuilt rows {rows} {cin}->{cout}: {us:.1f} us/launch ({2.0*rows*cin*cout/us/1e6:.1f} TF/s)")
