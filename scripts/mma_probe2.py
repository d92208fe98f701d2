"""Does a second issuing thread hide the tcgen05.commit drain? ns per MMA (total MMAs of both threads)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mfr_b200
from mfr_b200 import lib as mfr
lib = mfr.load()
torch.zeros(1, device="cuda")
for BN in (128, 208, 256):
    for mode, name, nthr in ((0, "one issuer", 1), (2, "two issuers, same accumulator", 2), (3, "two issuers, two accumulators", 2)):
        row = []
        for n in (1, 2, 4, 8, 12, 24):
            ns = lib.mfr_debug_mma_probe(n, BN, 1, 0, 2000, 148, mode)
            row.append(f"{n}:{ns / nthr:6.1f}")
        print(f"bf16 BN={BN} {name:32s} ns/MMA by MMAs-per-commit  " + "  ".join(row), flush=True)
