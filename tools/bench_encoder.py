"""Times the FFConvLSTM encoder on one GPU: whole forward and the LSTM recurrence alone.

    python tools/bench_encoder.py [--size recipe|default] [--tracks 6] [--frames 2000]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ensemble_svs_with_interactions_b200 import ops  # noqa: E402
from ensemble_svs_with_interactions_b200.model import FFConvLSTM  # noqa: E402

SIZES = {  # recipe: conf/train_acoustic/model/multitrack_acoustic_nnsvs_world_multi_ar_f0_diff_mgcbap.yaml:104-116
    "recipe": dict(in_dim=87, ff_hidden_dim=512, conv_hidden_dim=256, lstm_hidden_dim=128, out_dim=256, in_ph_start_idx=3,
                   in_ph_end_idx=50, embed_dim=256),
    "recipe_bap": dict(in_dim=87, ff_hidden_dim=256, conv_hidden_dim=128, lstm_hidden_dim=64, out_dim=128, in_ph_start_idx=3,
                       in_ph_end_idx=50, embed_dim=256),
    "default": dict(in_dim=87),  # model.py:803-806: 2048 / 1024 / 256
    # the stream models of the recipe's default (non-diffusion) config, multitrack_acoustic_nnsvs_world_multi_ar_f0.yaml:106-145
    "stream_mgc": dict(in_dim=1026, ff_hidden_dim=1024, conv_hidden_dim=512, lstm_hidden_dim=256, out_dim=60),
    "stream_bap": dict(in_dim=1026, ff_hidden_dim=256, conv_hidden_dim=128, lstm_hidden_dim=62, out_dim=5),
}


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="recipe", choices=sorted(SIZES))
    ap.add_argument("--tracks", type=int, default=6)
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--lstm-sweep", action="store_true", help="recurrence kernel alone over hidden sizes, with torch's cuDNN LSTM beside it")
    a = ap.parse_args()
    if a.lstm_sweep:
        B, T = a.tracks, a.frames
        for H in (32, 64, 128, 256):
            pre = torch.randn(B, T, 8 * H, device="cuda")
            w_hh = torch.randn(2, 4 * H, H, device="cuda") / H ** 0.5
            hb = torch.empty(B, T, 2 * H, device="cuda", dtype=torch.bfloat16)
            ms = timed(lambda: ops.lstm_f32(pre, w_hh, None, H, pre_layout="ntc", h_bf16=hb))
            ref = torch.nn.LSTM(2 * H, H, 1, bidirectional=True, batch_first=True).cuda().eval()   # the reference's own GPU path (cuDNN)
            xin = torch.randn(B, T, 2 * H, device="cuda")
            with torch.no_grad():
                ms_ref = timed(lambda: ref(xin))
            print(json.dumps({"H": H, "tracks": B, "frames": T, "svsk_lstm_ms": round(ms, 3), "us_per_step": round(ms * 1e3 / T, 3),
                              "cudnn_lstm_layer_ms_incl_input_gemm": round(ms_ref, 3)}))
        return
    cfg = SIZES[a.size]
    torch.manual_seed(0)
    B, T = a.tracks, a.frames
    x = torch.randn(B, T, cfg["in_dim"], device="cuda")
    if cfg.get("embed_dim"):
        x[..., 3:50] = torch.nn.functional.one_hot(torch.randint(0, 47, (B, T), device="cuda"), 47).float()
    out = {"size": a.size, "tracks": B, "frames": T}
    for prec in ("bf16", "fp32"):
        m = FFConvLSTM(**cfg, precision=prec).cuda().eval()
        out[f"forward_ms_{prec}"] = round(timed(lambda: m(x, [T] * B)), 3)
    H = m.padded_hidden
    pre = torch.randn(B, T, 8 * H, device="cuda")
    w_hh = torch.randn(2, 4 * H, H, device="cuda") / H ** 0.5
    hb = torch.empty(B, T, 2 * H, device="cuda", dtype=torch.bfloat16)
    ms = timed(lambda: ops.lstm_f32(pre, w_hh, None, H, pre_layout="ntc", h_bf16=hb))
    out["lstm_layer_ms"] = round(ms, 3)
    out["lstm_us_per_step"] = round(ms * 1e3 / T, 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
