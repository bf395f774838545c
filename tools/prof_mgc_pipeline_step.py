"""A few DDPM steps of the pipeline's mgc model at 6 tracks x 6000 frames (config 4 shape), eager launches — the command
profiled for profiles/r01z_launches_mgc_6x6000.*"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ensemble_svs_with_interactions_b200.diffsinger import DiffNet, GaussianDiffusion  # noqa: E402

torch.manual_seed(0)
m = GaussianDiffusion(256, 60, DiffNet(60, 256, 20, 256, 4), K_step=4).cuda().eval()
with torch.no_grad():
    m.denoise_fn.output_projection.weight.normal_(0, 0.02)
m.use_cuda_graph = False
cond = torch.randn(6, 6000, 256, device="cuda")
with torch.no_grad():
    y = m.inference(cond)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
