"""Per-slot timeline of one layer of the two-tiles-per-pair residual stack (diffnet_stack_duo_sm100.cu, C = 128) from its
clock64 stamps (SVSK_DIFFNET_TIMELINE; profiling aid).  usage: python tools/timeline_stack_duo.py [B=6] [T=6000]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ensemble_svs_with_interactions_b200 import ops  # noqa: E402
from ensemble_svs_with_interactions_b200.diffsinger import DiffNet, GaussianDiffusion  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
os.environ["SVSK_STACK_DUO"] = "1"
torch.manual_seed(0)
m = GaussianDiffusion(128, 5, DiffNet(5, 128, 10, 128, 4), K_step=100).to("cuda").eval()
plan = m.denoise_fn.bf16_plan()
table = m._step_table()
cond = torch.randn(B, T, plan.H, device="cuda").to(torch.bfloat16)
xb0 = torch.randn(B, T, plan.C, device="cuda").to(torch.bfloat16)
e0, e1 = torch.empty_like(xb0), torch.empty_like(xb0)
skip = torch.empty(B, T, plan.C, device="cuda")
flags = torch.empty((B * 2 * ((T + 255) // 256),), device="cuda", dtype=torch.int32)


def launch():
    ops.diffnet_stack_bf16(xb0, e0, e1, skip, cond, plan.w1p_all, plan.woutp_all, table[:, 50:51], plan.bout_all, flags,
                           plan.dilations, stepbias_batch_stride=0, stepbias_layer_stride=table.stride(0))


for _ in range(3):
    launch()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    launch()
b.record(); b.synchronize()
us = a.elapsed_time(b) * 1e3 / 20
print(f"B={B} T={T}: {us:.1f} us per launch of {plan.L} layers")
n_cta = B * 2 * (((T + 255) // 256 + 1) // 2)
dbg = torch.zeros(n_cta * 64, dtype=torch.int64, device="cuda")
os.environ["SVSK_DIFFNET_TIMELINE"] = str(dbg.data_ptr())
launch()
torch.cuda.synchronize()
os.environ.pop("SVSK_DIFFNET_TIMELINE")
d = dbg.view(n_cta, 64).cpu()
lead = d[0::2]                      # leader CTAs (MMA thread stamps live there)
lead = lead[lead[:, 40] > 0]        # both slots live
names = {0: "xe_ready seen (act)", 1: "edge rows stored", 2: "flag published", 3: "neighbours' flags seen", 4: "fences done",
         5: "G buffer free", 8: "MMA: centre+acc ready", 9: "MMA: halo landed", 10: "MMA: GEMM1 issued", 11: "MMA: G ready",
         12: "MMA: GEMM2 issued", 16: "epi: D1 complete", 17: "epi: gated", 18: "epi: D2 complete", 19: "epi: residual written",
         20: "epi: skip handed over"}
t0 = lead[:, 8:9]                   # slot 0: MMA thread sees centre rows + accumulator of the stamped layer
for u in (0, 1):
    rel = (lead[:, 32 * u:32 * u + 32] - t0).float().median(dim=0).values
    print(f"slot {u} (median over {lead.shape[0]} leader CTAs, cycles since slot 0's 'centre+acc ready' of the stamped layer): "
          + " | ".join(f"{n} @{int(rel[i])}" for i, n in sorted(names.items())))
print(f"whole kernel: {int((d[:, 62] - d[:, 63]).float().median())} cycles = {int((d[:, 62] - d[:, 63]).float().median()) / plan.L:.0f} per layer")
