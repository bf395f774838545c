"""Timing and per-layer timeline of the one-launch residual stack at the BASELINE config-2 shape (profiling aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ensemble_svs_with_interactions_b200 import ops  # noqa: E402

B, T = bench.B, bench.T
if len(sys.argv) > 2:
    B, T = int(sys.argv[1]), int(sys.argv[2])
if os.environ.get("SVSK_STACK_BENCH_MODEL"):   # "M,H,L,C", e.g. 5,128,10,128 = the pipeline's bap model
    from ensemble_svs_with_interactions_b200.diffsinger import DiffNet, GaussianDiffusion
    M_, H_, L_, C_ = (int(v) for v in os.environ["SVSK_STACK_BENCH_MODEL"].split(","))
    torch.manual_seed(0)
    m = GaussianDiffusion(H_, M_, DiffNet(M_, H_, L_, C_, 4), K_step=100).to("cuda").eval()
else:
    m = bench.build_model().to("cuda")
plan = m.denoise_fn.bf16_plan()
cond = torch.randn(B, T, plan.H, device="cuda").to(torch.bfloat16)
table = m._step_table()
xb0 = torch.randn(B, T, plan.C, device="cuda").to(torch.bfloat16)
e0b, e1b = torch.empty_like(xb0), torch.empty_like(xb0)
skip32 = torch.zeros(B, T, plan.C, device="cuda")
flags = torch.empty((B * 2 * ((T + 255) // 256),), device="cuda", dtype=torch.int32)
assert ops.diffnet_stack_fits(B, T, plan.C, plan.H)
# SVSK_STACK_BENCH_PCOND=1: with the conditioner projection precomputed (what a sampling run does), else inside the GEMM
pcond = None
if os.environ.get("SVSK_STACK_BENCH_PCOND"):
    ta, tb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m.denoise_fn.cond_projection_bf16(cond, plan)
    ta.record()
    pcond = m.denoise_fn.cond_projection_bf16(cond, plan)
    tb.record(); tb.synchronize()
    assert pcond is not None
    print(f"conditioner projection of all layers, once per run: {ta.elapsed_time(tb) * 1e3:.0f} us", flush=True)


def launch():
    ops.diffnet_stack_bf16(xb0, e0b, e1b, skip32, cond, plan.w1p_all, plan.woutp_all, table[:, 50:51], plan.bout_all, flags,
                           plan.dilations, stepbias_batch_stride=0, stepbias_layer_stride=table.stride(0), pcond=pcond)


for _ in range(3):
    launch()
torch.cuda.synchronize()
if os.environ.get("SVSK_PROFILE_RANGE"):
    torch.cuda.profiler.start()
    launch()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    launch()
b.record(); b.synchronize()
us = a.elapsed_time(b) * 1e3 / 20
flops = 2.0 * B * T * (2 * plan.C * (3 * plan.C + plan.H) + 2 * plan.C * plan.C) * plan.L
print(f"B={B} T={T}: stack of {plan.L} blocks {us:8.1f} us/launch = {us / plan.L:6.2f} us/layer, {flops / us / 1e6:7.1f} TFLOP/s", flush=True)

dbg = torch.zeros(512 * 32, dtype=torch.int64, device="cuda")
os.environ["SVSK_DIFFNET_TIMELINE"] = str(dbg.data_ptr())
launch()
torch.cuda.synchronize()
os.environ.pop("SVSK_DIFFNET_TIMELINE")
d = dbg.view(512, 32).cpu()
d = d[(d[:, 15] > 0) & (d[:, 2] > 0)]
d[:, 18:] = torch.where(d[:, 18:] == 0, d[:, :1], d[:, 18:])   # stamps a mode does not take
d[:, 1:16] = torch.where(d[:, 1:16] == 0, d[:, :1], d[:, 1:16])
names = {24: "epilogue: D2[0] complete", 25: "residual written", 26: "D2[1] complete", 29: "8 skip slabs staged", 30: "G buffer released", 28: "edge thread (row 0) past halo_free", 9: "edge thread residual written",
         10: "last epilogue warp residual written", 27: "layer 2: xc_ready seen (MMA thread)", 11: "flag published", 12: "neighbours' flags seen",
         13: "halo loads issued", 14: "G buffer free", 2: "layer 1: centre rows ready (MMA thread)", 3: "halo rows landed", 4: "first cond tile + D1[1] drained", 5: "block 0 issued",
         6: "block 1 issued", 7: "G ready", 8: "GEMM2 issued", 20: "gating 0 starts", 21: "gating 0 ends", 22: "gating 1 starts",
         23: "gating 1 ends", 31: "all epilogue threads through layer 1", 18: "layer 2: centre rows ready (MMA thread)",
         19: "layer 2: first weight tile landed", 15: "kernel end"}
rel = (d - d[:, :1]).float().median(dim=0).values
print(f"{d.shape[0]} leader CTAs; median cycles since CTA start: " + " | ".join(f"{n} @{int(rel[i])}" for i, n in names.items()))
raw = d.float().median(dim=0).values
print(f"   MMA thread: {int(raw[17])} of {plan.L * 40} ring entries not yet landed when reached, {int(raw[16])} cycles waiting for them")
print(f"   whole kernel {int(rel[15])} cycles = {int(rel[15]) / plan.L:.0f} per layer")
