"""Ablation timing of the fused uSFGAN block kernel (profiling aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ensemble_svs_with_interactions_b200 import ops  # noqa: E402

B, T = 6, 720000
g = torch.Generator().manual_seed(0)
xb = torch.randn(B, T, 64, device="cuda").to(torch.bfloat16)
out = torch.empty_like(xb)
auxb = torch.randn(B, T, 80, device="cuda").to(torch.bfloat16)
w1p, woutp = ops.usfgan_pack_block(torch.randn(128, 64, 3, device="cuda") * 0.05, torch.randn(128, 80, device="cuda") * 0.05,
                                   torch.randn(64, 64, device="cuda") * 0.1)
b1 = torch.zeros(128, device="cuda"); bo = torch.zeros(64, device="cuda")
d = torch.empty(B, 1, T, device="cuda").uniform_(2, 40)
idx = ops.pd_index(d, 4)
for name, kw in (("fixed d=8", dict(dilation=8)), ("adaptive", dict(idx=idx))):
    for ab in (0, 256, 0, 256, 1, 4, 5):
        os.environ["SVSK_USFGAN_ABLATE"] = str(ab)
        for _ in range(2):
            ops.usfgan_block_bf16(xb, out, auxb, w1p, woutp, b1, bo, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.usfgan_block_bf16(xb, out, auxb, w1p, woutp, b1, bo, **kw)
        e1.record(); e1.synchronize()
        us = e0.elapsed_time(e1) / 5 * 1e3
        print(f"{name:10s} ablate={ab:2d} (1=no epilogue, 4=no MMAs, 256=synchronous probe): {us:7.1f} us  "
              f"-> {us * 1e-6 * 1.85e9 / (B * ((T + 127) // 128) / 148):6.0f} cycles/tile", flush=True)

# role accounting (cycles per tile, averaged over CTAs)
names = ["prod wait empty", "mma wait operands", "mma wait G", "mma loop total", "tiles", "epi wait D1", "epi gating",
         "epi wait D2", "epi residual", "epi sync/store", "mma fence_after", "mma commit/arrive", "mma probe", "mma issue"]
for ab, name, kw in ((0, "fixed d=8", dict(dilation=8)), (0, "adaptive", dict(idx=idx)), (4, "fixed, no MMAs", dict(dilation=8)),
                     (5, "fixed, no MMAs/epilogue", dict(dilation=8))):
    os.environ["SVSK_USFGAN_ABLATE"] = str(ab)
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    os.environ["SVSK_USFGAN_TIMELINE"] = str(dbg.data_ptr())
    ops.usfgan_block_bf16(xb, out, auxb, w1p, woutp, b1, bo, **kw)
    torch.cuda.synchronize()
    os.environ.pop("SVSK_USFGAN_TIMELINE")
    dd = dbg.view(148, 16).float()
    tiles = dd[:, 4].clamp_min(1)
    print(name, " | ".join(f"{n}={float((dd[:, i] / tiles).mean()):.0f}" for i, n in enumerate(names) if i != 4))
