"""Times the fused encoder-layer tail kernel (csrc/encoder_tail.cu) at the benchmark's shapes:
coarse level C = 256, 2 x 16 x 6120 rows; fine level C = 128, 2 x 35 k x 25 rows. Prints us per launch and the
achieved TFLOP/s (14 C^2 flop per row)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mfr_b200
from mfr_b200 import lib as mfr
lib = mfr.load()
for C, rows in ((256, 2 * 16 * 6120), (128, 2 * 35000 * 25)):
    g = torch.Generator().manual_seed(0)
    msg = torch.randn(rows, C, generator=g).cuda().to(torch.bfloat16)
    x = torch.randn(rows, 2 * C, generator=g).cuda()
    xb = x.to(torch.bfloat16)
    w = [(torch.randn(o, i, generator=g) / i ** 0.5).cuda().to(torch.bfloat16) for o, i in ((C, C), (2 * C, 2 * C), (C, 2 * C))]
    nrm = [torch.ones(C).cuda(), torch.zeros(C).cuda(), torch.ones(C).cuda() * 0.01, torch.zeros(C).cuda()]
    args = [mfr.ptr(msg), mfr.ptr(xb), mfr.ptr(x)] + [mfr.ptr(t) for t in w] + [mfr.ptr(t) for t in nrm]
    mfr.check(lib.mfr_encoder_tail(*args, C, 1, rows, None, 3, mfr.cur_stream()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    mfr.check(lib.mfr_encoder_tail(*args, C, 1, rows, None, n, mfr.cur_stream()))
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"C={C} rows={rows}: {us:.1f} us per launch, {14.0 * C * C * rows / us / 1e6:.1f} TFLOP/s, {us * 148 / ((rows + 127) // 128):.2f} us per 128-row tile per SM", flush=True)
