"""Profiling target: one ParallelHn-uSFGAN forward (BASELINE config 3, 6 tracks x 30 s) inside a cudaProfiler range.
usage: ncu --profile-from-start off --metrics gpu__time_duration.sum ... python tools/prof_usfgan.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ensemble_svs_with_interactions_b200.usfgan.models import ParallelHnUSFGANGenerator  # noqa: E402

B, Fr, HOP, FS = 6, 6000, 120, 24000
T = Fr * HOP
pe = {"conv_layers": 3, "kernel_size": 5, "dilation": 1, "padding_mode": "replicate"}
torch.manual_seed(1234)
m = ParallelHnUSFGANGenerator(periodicity_estimator_params=pe).eval()
m.remove_weight_norm()
m = m.cuda()
g = torch.Generator().manual_seed(1)
c = torch.randn(B, 80, Fr + 4, generator=g).cuda()
f0 = torch.empty(B, 1, Fr).uniform_(110, 880, generator=g)
d = (FS / (f0 * 4)).repeat_interleave(HOP, dim=-1).cuda()
x = (torch.randn(B, 2, T, generator=g) * 0.1).cuda()
m(x, c, d, wave_only=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
y = m(x, c, d, wave_only=True)[0]
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", tuple(y.shape), bool(torch.isfinite(y).all()))
