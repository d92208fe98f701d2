"""Tensor-pipe rate by MMA shape: ns per tcgen05.mma (K = 32 bytes) for single CTAs (M = 128) and CTA pairs (cta_group::2,
M = 256 over two SMs), N from 64 to 256, 32 MMAs per commit, all SMs busy. Operands resident in shared memory."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfr_b200  # noqa: F401
from mfr_b200 import lib as mfr
lib = mfr.load()
for bf16 in (1, 0):
    for mode, name in ((0, "single M=128"), (4, "pair   M=256")):
        row = []
        for BN in (64, 96, 128, 160, 192, 208, 224, 256):
            ns = lib.mfr_debug_mma_probe(32, BN, bf16, 0, 2000, 148, mode)
            row.append(f"{BN}: {ns:6.1f}")
        print(f"{'bf16' if bf16 else 'tf32'} {name}  ns/MMA by N  " + "  ".join(row), flush=True)
