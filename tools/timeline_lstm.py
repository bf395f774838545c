"""SVSK_LSTM_TIMELINE=1 python tools/timeline_lstm.py — prints the per-step clock stamps of the LSTM recurrence kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ensemble_svs_with_interactions_b200 import ops  # noqa: E402

for H in (32, 64, 128, 256):
    pre = torch.randn(6, 400, 8 * H, device="cuda")
    w_hh = torch.randn(2, 4 * H, H, device="cuda") / H ** 0.5
    hb = torch.empty(6, 400, 2 * H, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.lstm_f32(pre, w_hh, None, H, pre_layout="ntc", h_bf16=hb)
    torch.cuda.synchronize()
