"""Where the token GEMMs (q|k|v and friends, bf16 in / bf16 out) spend their time: one instrumented launch per shape prints the
cycles block 0's producer / MMA issuer / first epilogue warp wait on their barriers, then the timed launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mfr_b200  # noqa: F401
from mfr_b200 import lib as mfr
lib = mfr.load()
for (N, rows, cin, cout) in ((32, 6120, 256, 768), (32, 6120, 256, 256), (32, 6120, 256, 512), (2, 660000, 128, 384)):
    x = torch.randn(N, 1, rows, cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(cout, 1, cin, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(N, 1, rows, cout, device="cuda", dtype=torch.bfloat16)
    us = lib.mfr_debug_time_conv(mfr.ptr(x), N, 1, rows, cin, mfr.ptr(w), cout, 1, 1, mfr.ptr(out), 20, 3, None)
    gb = N * rows * (cin + cout) * 2 / 1e9
    print(f"{(N, rows, cin, cout)}: {us:.1f} us/launch, {gb / us * 1e6:.0f} GB/s of operand + result bytes", flush=True)

# the fine-level q|k|v GEMM as the engine runs it: A rows are the x half of the [x | msg] token rows (row stride 2C)
def timed(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
N, rows, cin, cout = 2, 660000, 128, 384
w = (torch.randn(cout, 1, cin, device="cuda") * 0.05).to(torch.bfloat16)
out = torch.empty(N, 1, rows, cout, device="cuda", dtype=torch.bfloat16)
for ldx in (128, 256):
    x = torch.randn(N, 1, rows, ldx, device="cuda").to(torch.bfloat16)
    us = timed(lambda: mfr.check(lib.mfr_conv2d_nhwc_bf16(mfr.ptr(x), N, 1, rows, cin, ldx, mfr.ptr(w), cin, cout, 1, 1, 1, 0, None, None, cout,
                                                          3, 1.0, mfr.ptr(out), cout, 1, mfr.cur_stream())))
    print(f"fine q|k|v shape, A row stride {ldx}: {us:.1f} us/launch", flush=True)
