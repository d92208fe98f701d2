"""CTA pairs (tcgen05 cta_group::2) vs single CTAs on the 3x3 backbone layers at the bench batch (32 images, bf16 maps,
196-channel maps stored with a channel stride of 200). One process per setting (the tiling is chosen at build)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    import mfr_b200  # noqa: F401
    from mfr_b200 import lib as mfr
    lib = mfr.load()
    N = int(os.environ.get("N", 32))
    SHAPES = ((360, 272, 128, 128, True), (360, 272, 128, 128, False), (360, 272, 196, 196, False), (360, 272, 196, 128, False),
              (180, 136, 196, 196, True), (180, 136, 256, 256, False), (180, 136, 256, 196, False), (90, 68, 256, 256, True))
    for (H, W, cin, cout, use_res) in SHAPES:
        ci, co = (cin + 7) // 8 * 8, (cout + 7) // 8 * 8
        x = torch.randn(N, H, W, ci, device="cuda").to(torch.bfloat16)
        w = (torch.randn(cout, 9, ci, device="cuda") * 0.05).to(torch.bfloat16)
        out = torch.empty(N, H, W, co, device="cuda", dtype=torch.bfloat16)
        res = torch.randn(N, H, W, co, device="cuda").to(torch.bfloat16) if use_res else None
        us = lib.mfr_debug_time_conv(mfr.ptr(x), N, H, W, cin, mfr.ptr(w), cout, 3, 1, mfr.ptr(out), 20, 7, mfr.ptr(res) if use_res else None)
        fl = 2.0 * N * H * W * cout * cin * 9
        print(f"{(N, H, W, cin, cout)} res={int(use_res)}: {us:.1f} us/launch  {fl / us / 1e6:.0f} TF/s", flush=True)
else:
    for pair in ("0", "1"):
        print("MFR_CONV_PAIR=" + pair, flush=True)
        env = dict(os.environ, MFR_CONV_PAIR=pair, MFR_CONV_DEBUG="1")
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, check=False, timeout=300)
