"""SVSK_STEP_TIMELINE=1 python tools/timeline_step.py — clock stamps of one CTA of the DDPM step kernel at config 2."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

m = bench.build_model().to("cuda")
cond = torch.randn(bench.B, bench.T, 256, device="cuda")
m.use_cuda_graph = False   # the stamps are read back after every launch
with torch.no_grad():
    y = m.inference(cond)
torch.cuda.synchronize()
