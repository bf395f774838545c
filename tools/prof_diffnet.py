"""Profiling target: eager (no CUDA graph) DDPM steps of the BASELINE config-2 denoiser so that ncu sees every
launch of a steady-state step.  One untimed pass first (weight packing, step-bias table, allocator warm-up), then the
profiled range.  usage: ncu --profile-from-start off ... python tools/prof_diffnet.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
m = bench.build_model().to("cuda")
m.use_cuda_graph = False
m.K_step = steps  # only the first `steps` entries of the schedule tables are used
g = torch.Generator().manual_seed(0)
cond = torch.randn(bench.B, bench.T, 256, generator=g).cuda()
m.inference(cond)
torch.cuda.synchronize()
torch.cuda.profiler.start()
y = m.inference(cond)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", tuple(y.shape), float(y.abs().mean()))
