"""One bf16 forward of the recipe-size FFConvLSTM encoder (6 tracks x 2000 frames) — the command profiled with ncu for
profiles/r01z_encoder_*.  python tools/prof_encoder.py [default]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ensemble_svs_with_interactions_b200.model import FFConvLSTM  # noqa: E402
from tools.bench_encoder import SIZES  # noqa: E402

cfg = SIZES[sys.argv[1] if len(sys.argv) > 1 else "recipe"]
torch.manual_seed(0)
m = FFConvLSTM(**cfg, precision="bf16").cuda().eval()
x = torch.randn(6, 2000, cfg["in_dim"], device="cuda")
if cfg.get("embed_dim"):
    x[..., 3:50] = torch.nn.functional.one_hot(torch.randint(0, 47, (6, 2000), device="cuda"), 47).float()
for _ in range(2):
    y = m(x, [2000] * 6)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.abs().mean()))
