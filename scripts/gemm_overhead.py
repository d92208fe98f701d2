"""Fixed cost of one tcgen05 GEMM launch (empty / tiny problems, back to back, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mfr_b200
from mfr_b200 import lib as mfr
lib = mfr.load()
for rows in (128, 1024, 6120, 12240, 6120 * 8):
    for (cin, cout) in ((256, 256), (512, 512)):
        x = torch.randn(1, 1, rows, cin, device="cuda"); w = torch.randn(cout, 1, cin, device="cuda") * 0.05
        out = torch.empty(1, 1, rows, cout, device="cuda")
        def run():
            mfr.check(lib.mfr_conv2d_nhwc_tf32(mfr.ptr(x), 1, 1, rows, cin, cin, mfr.ptr(w), cout, 1, 1, 1, 0, 0, None, None, cout, 0, 1.0, mfr.ptr(out), cout, mfr.cur_stream()))
        for _ in range(5): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        print(f"rows {rows} {cin}->{cout}: {us:.1f} us/launch ({2.0*rows*cin*cout/us/1e6:.1f} TF/s)")
