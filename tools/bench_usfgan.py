"""BASELINE config 3: ParallelHn-uSFGAN (recipe config: 20A/4 + 5F/5 + 30F/3, 64/128/64, aux 80, hop 120) for
6 tracks x 30 s at 24 kHz, with a per-component breakdown (CUDA events)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ensemble_svs_with_interactions_b200 import _lib, ops  # noqa: E402
from ensemble_svs_with_interactions_b200.usfgan.models import ParallelHnUSFGANGenerator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
SECONDS = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
FS, HOP = 24000, 120
Fr = int(SECONDS * FS / HOP)
T = Fr * HOP
pe = {"conv_layers": 3, "kernel_size": 5, "dilation": 1, "padding_mode": "replicate"}
torch.manual_seed(1234)
m = ParallelHnUSFGANGenerator(periodicity_estimator_params=pe).eval()
with torch.no_grad():
    m.periodicity_estimator.layers[-2].weight_v.normal_(0, 0.05)
m.remove_weight_norm()
m = m.cuda()
g = torch.Generator().manual_seed(1)
c = torch.randn(B, 80, Fr + 4, generator=g).cuda()
f0 = torch.empty(B, 1, Fr).uniform_(110, 880, generator=g)
d = (FS / (f0 * 4)).repeat_interleave(HOP, dim=-1).cuda()
x = (torch.randn(B, 2, T, generator=g) * 0.1).cuda()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps, out


for prec in ("bf16",) + (("fp32",) if "--fp32" in sys.argv else ()):
    m.precision = prec
    n0 = _lib.launch_count
    ms, out = timed(lambda: m(x, c, d, wave_only=True)[0])
    print(f"[{prec}] ParallelHn forward B={B} x {SECONDS:.0f}s: {ms:.1f} ms -> {B * SECONDS / (ms / 1e3):.0f} audio-sec/sec "
          f"({(_lib.launch_count - n0) // 4} libsvsk launches/pass), finite={bool(torch.isfinite(out).all())}", flush=True)

# breakdown (bf16)
m.precision = "bf16"
ms_up, cu = timed(lambda: m.upsample_net(c))
ms_pe, a = timed(lambda: m.periodicity_estimator(cu))
auxb, _ = ops.nct_to_ntc(cu, Cp=80)
hb = torch.randn(B, T, 64, device="cuda").to(torch.bfloat16)
cache = {}
ms_h, _ = timed(lambda: m.harmonic_network.forward_ntc_bf16(hb, auxb, d, cache))
ms_n, _ = timed(lambda: m.noise_network.forward_ntc_bf16(hb, auxb, d, cache))
ms_f, yb = timed(lambda: m.filter_network.forward_ntc_bf16(hb, auxb, d, cache))
y = ops.ntc_bf16_to_nct_f32(yb, 64)
ms_last, _ = timed(lambda: m._conv_last(y))
ms_cv, _ = timed(lambda: ops.nct_to_ntc(y))
ms_cv2, _ = timed(lambda: ops.ntc_bf16_to_nct_f32(yb, 64))
print(f"breakdown ms: upsample {ms_up:.2f} | periodicity {ms_pe:.2f} | harmonic(20A) {ms_h:.2f} | noise(5F) {ms_n:.2f} | "
      f"filter(30F) {ms_f:.2f} | conv_last {ms_last:.2f} | nct->ntc {ms_cv:.2f} | ntc->nct {ms_cv2:.2f}")
flop_block = 2.0 * B * T * 38912
print(f"per block: adaptive {ms_h / 20 * 1e3:.0f} us, fixed(filter) {ms_f / 30 * 1e3:.0f} us -> "
      f"{flop_block / (ms_f / 30 * 1e-3) / 1e12:.0f} TFLOP/s, {B * T * 416 / (ms_f / 30 * 1e-3) / 1e9:.0f} GB/s algorithmic")
