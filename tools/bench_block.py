"""Micro-benchmark of the fused DiffNet block kernel at the BASELINE config-2 shape for several time tiles."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ensemble_svs_with_interactions_b200 import ops  # noqa: E402

B, T = bench.B, bench.T
if len(sys.argv) > 2:
    B, T = int(sys.argv[1]), int(sys.argv[2])
m = bench.build_model().to("cuda")
den = m.denoise_fn
plan = den.bf16_plan()
cond = torch.randn(B, T, 256, device="cuda").to(torch.bfloat16)
table = m._step_table()
sb = [tl[50] for tl in table]
xb0 = torch.randn(B, T, plan.C, device="cuda").to(torch.bfloat16)
xb1 = torch.empty_like(xb0)
x32 = torch.randn(B, T, plan.C, device="cuda")
skip32 = torch.zeros(B, T, plan.C, device="cuda")
flops = 2.0 * B * T * bench.BLOCK_MAC_PER_FRAME
for tile, kernel in ((0, 3), (0, 2), (96, 1)):
    def blocks():
        cur, nxt = xb0, xb1
        for i, lw in enumerate(plan.layers):
            ops.diffnet_block_bf16(cur, nxt, x32, skip32, cond, lw["w1p"], lw["woutp"], sb[i], lw["bout"],
                                   dilation=lw["dilation"], stepbias_batch_stride=0, init_skip=(i == 0), write_x=True,
                                   time_tile=tile, kernel=kernel)
            cur, nxt = nxt, cur
    for _ in range(3):
        blocks()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        blocks()
    e1.record()
    e1.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 200
    print(f"B={B} T={T} kernel={kernel} tile={tile:3d}: {us:7.2f} us/launch  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
