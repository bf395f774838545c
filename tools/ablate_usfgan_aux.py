"""How much of a uSFGAN block is the aux stream?  Same block kernel with aux widths 80 / 16 (profiling aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ensemble_svs_with_interactions_b200 import ops
B, T = 6, 720000
xb = torch.randn(B, T, 64, device="cuda").to(torch.bfloat16); out = torch.empty_like(xb)
d = torch.empty(B, 1, T, device="cuda").uniform_(2, 40); idx = ops.pd_index(d, 4)
for A in (80, 64, 16):
    auxb = torch.randn(B, T, A, device="cuda").to(torch.bfloat16)
    w1p, woutp = ops.usfgan_pack_block(torch.randn(128, 64, 3, device="cuda") * 0.05, torch.randn(128, A, device="cuda") * 0.05,
                                       torch.randn(64, 64, device="cuda") * 0.1)
    b1 = torch.zeros(128, device="cuda"); bo = torch.zeros(64, device="cuda")
    for name, kw in (("fixed d=8", dict(dilation=8)), ("adaptive", dict(idx=idx))):
        for _ in range(2):
            ops.usfgan_block_bf16(xb, out, auxb, w1p, woutp, b1, bo, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.usfgan_block_bf16(xb, out, auxb, w1p, woutp, b1, bo, **kw)
        e1.record(); e1.synchronize()
        print(f"aux width {A:3d} {name:10s}: {e0.elapsed_time(e1) / 5 * 1e3:7.1f} us", flush=True)
