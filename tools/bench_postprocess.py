"""Times the device post-processing (row f3) at the pipeline's shape: 6 tracks x 6000 frames, mgc 60 + bap 5 dims."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ensemble_svs_with_interactions_b200 import postprocess as pp  # noqa: E402
from tools.bench_encoder import timed  # noqa: E402

m = torch.randn(6, 6000, 60, device="cuda")
b = torch.randn(6, 6000, 5, device="cuda")
gv = torch.rand(60, device="cuda") + 0.5
mask = torch.rand(6, 6000, device="cuda") > 0.2
out = {"lowpass_mgc_ms": round(timed(lambda: pp.lowpass_filter(m, 200, cutoff=50)), 3),
       "lowpass_bap_ms": round(timed(lambda: pp.lowpass_filter(b, 200, cutoff=50)), 3),
       "variance_scaling_mgc_ms": round(timed(lambda: pp.variance_scaling(gv, m, 2, mask)), 3)}
print(json.dumps(out))
