"""BASELINE config 3, second half: the other uSFGAN generators at their defaults (6 tracks x 30 s @ 24 kHz, bf16):
USFGANGenerator (30 adaptive + 30 fixed blocks) and CascadeHnUSFGANGenerator, next to the recipe's ParallelHn."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ensemble_svs_with_interactions_b200 import _lib  # noqa: E402
from ensemble_svs_with_interactions_b200.usfgan.models import (CascadeHnUSFGANGenerator, ParallelHnUSFGANGenerator,  # noqa: E402
                                                              USFGANGenerator)

B, SECONDS, FS, HOP = 6, 30.0, 24000, 120
Fr = int(SECONDS * FS / HOP)
T = Fr * HOP
g = torch.Generator().manual_seed(1)
c = torch.randn(B, 80, Fr + 4, generator=g).cuda()
f0 = torch.empty(B, 1, Fr).uniform_(110, 880, generator=g)
d = (FS / (f0 * 4)).repeat_interleave(HOP, dim=-1).cuda()
pe = {"conv_layers": 3, "kernel_size": 5, "dilation": 1, "padding_mode": "replicate"}
for name, ctor, cin in (("USFGANGenerator (30A + 30F)", lambda: USFGANGenerator(), 1),
                        ("CascadeHnUSFGANGenerator", lambda: CascadeHnUSFGANGenerator(periodicity_estimator_params=pe), 2),
                        ("ParallelHnUSFGANGenerator (recipe)", lambda: ParallelHnUSFGANGenerator(periodicity_estimator_params=pe), 2)):
    torch.manual_seed(1234)
    m = ctor().eval()
    m.remove_weight_norm()
    m = m.cuda()
    x = (torch.randn(B, cin, T, generator=g) * 0.1).cuda()
    fn = (lambda: m(x, c, d, wave_only=True)[0]) if getattr(m, "supports_wave_only", False) else (lambda: m(x, c, d)[0])
    with torch.no_grad():
        fn(); fn()
        torch.cuda.synchronize()
        n0 = _lib.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = fn()
        e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{name}: precision {getattr(m, 'precision', '?')}/{m.resolved_precision() if hasattr(m, 'resolved_precision') else ''} "
          f"{ms:.1f} ms per pass -> {B * SECONDS / (ms / 1e3):.0f} audio-sec/s, {(_lib.launch_count - n0) // 3} libsvsk launches, "
          f"finite={bool(torch.isfinite(out).all())}", flush=True)
