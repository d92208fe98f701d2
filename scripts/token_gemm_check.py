"""Token-GEMM shapes of the coarse transformer at the bench batch (32 images x 6120 tokens): time per launch for the
four operand-type combinations (flags bit 0: bf16 inputs, bit 1: bf16 output) to separate operand traffic / MMA kind
from the epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mfr_b200  # noqa: F401
from mfr_b200 import lib as mfr
lib = mfr.load()
N, W = 32, 6120
for (cin, cout) in ((256, 768), (256, 256), (512, 512), (512, 256)):
    for flags in (0, 2, 1, 3):
        idt = torch.bfloat16 if flags & 1 else torch.float32
        odt = torch.bfloat16 if flags & 2 else torch.float32
        x = torch.randn(N, 1, W, cin, device="cuda").to(idt)
        w = (torch.randn(cout, 1, cin, device="cuda") * 0.05).to(idt)
        out = torch.empty(N, 1, W, cout, device="cuda", dtype=odt)
        us = lib.mfr_debug_time_conv(mfr.ptr(x), N, 1, W, cin, mfr.ptr(w), cout, 1, 1, mfr.ptr(out), 20, flags, None)
        fl = 2.0 * N * W * cout * cin
        by = N * W * (cin * x.element_size() + cout * out.element_size())
        print(f"{cin:4d}->{cout:4d} in={'bf16' if flags & 1 else 'fp32'} out={'bf16' if flags & 2 else 'fp32'}: {us:7.1f} us  "
              f"{fl / us / 1e6:6.0f} TF/s  {by / us / 1e6:5.2f} TB/s", flush=True)
