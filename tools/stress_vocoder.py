"""Race hunt for the uSFGAN block kernels INSIDE full vocoder passes (the two-epilogue-group variant
of round 1 failed intermittently only there): repeat whole ParallelHn / Cascade / plain generator forwards on fixed inputs —
recipe depth, frame-rate aux projection, adaptive + fixed blocks back to back on one stream — and require every waveform to
be bit-identical to the first one.  usage: python tools/stress_vocoder.py [repeats=300]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ensemble_svs_with_interactions_b200.usfgan.models import (CascadeHnUSFGANGenerator, ParallelHnUSFGANGenerator,  # noqa: E402
                                                               USFGANGenerator)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
FS, HOP = 24000, 120
pe = {"conv_layers": 3, "kernel_size": 5, "dilation": 1, "padding_mode": "replicate"}
bad = 0
cases = (("ParallelHn", ParallelHnUSFGANGenerator, dict(periodicity_estimator_params=pe), 6, 6000, n),
         ("ParallelHn", ParallelHnUSFGANGenerator, dict(periodicity_estimator_params=pe), 3, 1237, n),
         ("ParallelHn", ParallelHnUSFGANGenerator, dict(periodicity_estimator_params=pe), 1, 131, n),
         ("CascadeHn", CascadeHnUSFGANGenerator, dict(periodicity_estimator_params=pe), 2, 911, max(n // 3, 1)),
         ("uSFGAN", USFGANGenerator, dict(), 2, 911, max(n // 3, 1)))
for name, cls, kw, B, Fr, reps in cases:
    torch.manual_seed(1234)
    m = cls(**kw).eval()
    m.remove_weight_norm()
    m = m.cuda()
    T = Fr * HOP
    g = torch.Generator().manual_seed(B * Fr)
    c = torch.randn(B, 80, Fr + 4, generator=g).cuda()
    f0 = torch.empty(B, 1, Fr).uniform_(80, 1000, generator=g)
    d = (FS / (f0 * 4)).repeat_interleave(HOP, dim=-1).cuda()
    nch = 1 if cls is USFGANGenerator else 2
    x = (torch.randn(B, nch, T, generator=g) * 0.1).cuda()
    ref = None
    for it in range(reps):
        out = m(x, c, d, wave_only=True)[0] if cls is not USFGANGenerator else m(x, c, d)[0]
        if ref is None:
            ref = out.clone()
            assert m.resolved_precision() == "bf16" if hasattr(m, "resolved_precision") else True
        elif not torch.equal(out, ref):
            bad += 1
            print(f"{name} B={B} Fr={Fr}: pass {it} differs, max|d|={float((out - ref).abs().max()):.3e}", flush=True)
    torch.cuda.synchronize()
    print(f"{name} B={B} x {Fr} frames: {reps} passes, finite={bool(torch.isfinite(ref).all())}, "
          f"aux projection={getattr(m, 'aux_projection', '?')}", flush=True)
print("STRESS_VOCODER", "OK" if bad == 0 else f"FAILED ({bad} differing passes)")
sys.exit(1 if bad else 0)
