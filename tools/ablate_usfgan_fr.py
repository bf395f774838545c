"""Timing of the uSFGAN block kernel: sample-rate aux stream vs frame-rate aux projection, with the ablation flags and the
per-role cycle accounting of the frame-rate kernel (profiling aid; operands are random, only the shapes matter)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ensemble_svs_with_interactions_b200 import ops  # noqa: E402

B, T, HOP, REACH = 6, 720000, 120, 152
bf = torch.bfloat16
xb = torch.randn(B, T, 64, device="cuda").to(bf)
out = torch.empty_like(xb)
auxb = torch.randn(B, T, 80, device="cuda").to(bf)
wt, wa, wo = torch.randn(128, 64, 3, device="cuda") * 0.05, torch.randn(128, 80, device="cuda") * 0.05, torch.randn(64, 64, device="cuda") * 0.1
w1p_s, woutp = ops.usfgan_pack_block(wt, wa, wo)
w1p_f, _ = ops.usfgan_pack_block(wt, None, wo)
b1 = torch.zeros(128, device="cuda"); bo = torch.zeros(64, device="cuda")
Tf = T // HOP
# dilation factors as the wrapper makes them: constant over a hop (F0 is frame-level) -> most 8-row groups of a tile have
# consecutive source rows; "adaptive rnd" draws them per sample (no such group: every row is gathered by cp.async)
d = torch.empty(B, 1, Tf, device="cuda").uniform_(2, 40).repeat_interleave(HOP, dim=-1).contiguous()
idx = ops.pd_index(d, 4)
idx_rnd = ops.pd_index(torch.empty(B, 1, T, device="cuda").uniform_(2, 40), 4)
cinb = torch.randn(B, Tf, 80, device="cuda").to(bf)
w_all = torch.randn(3 * 128, 80, device="cuda").to(bf)
q, fpad = ops.usfgan_aux_frames(cinb, w_all, Tf, T, HOP, REACH)
u = (torch.rand((T + 127) // 128 * 128, 16, device="cuda") * 0.2).to(bf)
frames = ops.UsfganAuxFrames(u, q, fpad, HOP, REACH)
ntile = B * ((T + 127) // 128) / 148


def run(mode, **kw):
    if mode == "samples":
        ops.usfgan_block_bf16(xb, out, auxb, w1p_s, woutp, b1, bo, **kw)
    else:
        ops.usfgan_block_bf16(xb, out, None, w1p_f, woutp, b1, bo, frames=frames, frames_block=1, **kw)


def timed(mode, **kw):
    for _ in range(2):
        run(mode, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run(mode, **kw)
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / 5 * 1e3


os.environ.pop("SVSK_USFGAN_ABLATE", None)
for name, kw in (("fixed d=8", dict(dilation=8)), ("fixed d=512", dict(dilation=512)), ("adaptive", dict(idx=idx)),
                 ("adaptive rnd", dict(idx=idx_rnd))):
    for mode in ("samples", "frames", "samples", "frames"):
        us = timed(mode, **kw)
        print(f"{name:12s} {mode:8s}: {us:7.1f} us -> {us * 1e-6 * 1.85e9 / ntile:6.0f} cycles/tile", flush=True)
for name, kw in (("fixed d=8", dict(dilation=8)), ("adaptive", dict(idx=idx))):
    for ab in (1, 4, 5, 2, 8, 16, 2 + 8 + 16, 64):
        os.environ["SVSK_USFGAN_ABLATE"] = str(ab)
        us = timed("frames", **kw)
        print(f"{name:12s} frames ablate={ab} (1=no epilogue, 2=no MUFU, 4=no MMAs, 8=no st.global, 16=no LDTM, 64=no TMA tap groups): {us:7.1f} us -> {us * 1e-6 * 1.85e9 / ntile:6.0f} cycles/tile", flush=True)
    os.environ.pop("SVSK_USFGAN_ABLATE")

names = {0: "prod wait empty", 1: "mma wait operands", 2: "mma wait G", 3: "mma loop total", 13: "mma issue+commit",
         14: "gather wait empty", 15: "gather issue", 5: "epi0 wait D1", 6: "epi0 gating", 7: "epi0 wait D2",
         8: "epi0 residual+store", 9: "epi1 wait D1", 10: "epi1 gating", 11: "epi1 wait D2", 12: "epi1 residual+store"}
for name, kw in (("fixed d=8", dict(dilation=8)), ("adaptive", dict(idx=idx))):
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    os.environ["SVSK_USFGAN_TIMELINE"] = str(dbg.data_ptr())
    run("frames", **kw)
    torch.cuda.synchronize()
    os.environ.pop("SVSK_USFGAN_TIMELINE")
    dd = dbg.view(148, 16).float()
    tiles = dd[:, 4].clamp_min(1)
    print(name, "(cycles per tile, with clock64 accounting on):", " | ".join(f"{n}={float((dd[:, i] / tiles).mean()) * (2 if i in range(5, 13) else 1):.0f}" for i, n in names.items()),
          "(epilogue groups: per tile of their own)")
