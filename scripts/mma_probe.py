"""Tensor-pipe probe: ns per tcgen05.mma vs MMAs per commit (operands resident in smem, no TMA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mfr_b200
from mfr_b200 import lib as mfr
lib = mfr.load()
torch.zeros(1, device="cuda")
for grid in (1, 148):
    for bf16 in (1, 0):
        for BN in (128, 256):
            for wait_each in (0, 1):
                for two in (0, 1):
                    row = []
                    for n in (1, 2, 4, 8, 16, 32):
                        ns = lib.mfr_debug_mma_probe(n, BN, bf16, wait_each, 2000, grid, two)
                        row.append(f"{n}:{ns:7.1f}")
                    print(f"grid={grid:3d} bf16={bf16} BN={BN} wait_each={wait_each} two_commits={two}  ns/MMA by MMAs-per-commit  " + "  ".join(row), flush=True)
