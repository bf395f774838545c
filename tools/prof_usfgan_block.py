"""Profiling target: one fixed and one adaptive uSFGAN block launch (6 x 720 000 samples) inside a cudaProfiler range."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ensemble_svs_with_interactions_b200 import ops  # noqa: E402

B, T = 6, 720000
xb = torch.randn(B, T, 64, device="cuda").to(torch.bfloat16); out = torch.empty_like(xb)
auxb = torch.randn(B, T, 80, device="cuda").to(torch.bfloat16)
w1p, woutp = ops.usfgan_pack_block(torch.randn(128, 64, 3, device="cuda") * 0.05, torch.randn(128, 80, device="cuda") * 0.05,
                                   torch.randn(64, 64, device="cuda") * 0.1)
b1 = torch.zeros(128, device="cuda"); bo = torch.zeros(64, device="cuda")
idx = ops.pd_index(torch.empty(B, 1, T, device="cuda").uniform_(2, 40), 4)
for _ in range(2):
    ops.usfgan_block_bf16(xb, out, auxb, w1p, woutp, b1, bo, dilation=8)
    ops.usfgan_block_bf16(xb, out, auxb, w1p, woutp, b1, bo, idx=idx)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.usfgan_block_bf16(xb, out, auxb, w1p, woutp, b1, bo, dilation=8)
ops.usfgan_block_bf16(xb, out, auxb, w1p, woutp, b1, bo, idx=idx)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
