"""Three launches of the frame-rate-aux uSFGAN block kernel at config 3's shape (6 x 30 s @ 24 kHz) for ncu:
ncu --set full --import-source on -k regex:usfgan_block_fr -c 3 python tools/prof_usfgan_block_fr.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ensemble_svs_with_interactions_b200 import ops  # noqa: E402

B, T, HOP, REACH = 6, 720000, 120, 152
bf = torch.bfloat16
xb = torch.randn(B, T, 64, device="cuda").to(bf)
out = torch.empty_like(xb)
wt, wo = torch.randn(128, 64, 3, device="cuda") * 0.05, torch.randn(64, 64, device="cuda") * 0.1
w1p, woutp = ops.usfgan_pack_block(wt, None, wo)
b1 = torch.zeros(128, device="cuda"); bo = torch.zeros(64, device="cuda")
# dilation factors constant over a hop, as USFGANWrapper makes them (most 8-row tap groups then travel as TMA boxes)
idx = ops.pd_index(torch.empty(B, 1, T // HOP, device="cuda").uniform_(2, 40).repeat_interleave(HOP, dim=-1).contiguous(), 4)
Tf = T // HOP
q, fpad = ops.usfgan_aux_frames(torch.randn(B, Tf, 80, device="cuda").to(bf), torch.randn(128, 80, device="cuda").to(bf), Tf, T, HOP, REACH)
frames = ops.UsfganAuxFrames((torch.rand((T + 127) // 128 * 128, 16, device="cuda") * 0.2).to(bf), q, fpad, HOP, REACH)
for kw in (dict(dilation=8), dict(dilation=8), dict(idx=idx)):
    ops.usfgan_block_bf16(xb, out, None, w1p, woutp, b1, bo, frames=frames, **kw)
torch.cuda.synchronize()
print("done")
