"""Race hunt (debug aid): repeat the persistent kernels many times on fixed inputs; every output must be bit-identical."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ensemble_svs_with_interactions_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
m = bench.build_model().to("cuda"); den = m.denoise_fn; plan = den.bf16_plan(); table = m._step_table()
bad = 0
for B, T in ((6, 2000), (3, 6000), (5, 517), (2, 2049)):
    g = torch.Generator().manual_seed(B * T)
    cond = torch.randn(B, T, plan.H, generator=g).cuda().to(torch.bfloat16)
    xb0 = torch.randn(B, T, plan.C, generator=g).cuda().to(torch.bfloat16)
    e0, e1 = torch.empty_like(xb0), torch.empty_like(xb0)
    flags = torch.empty((B * 2 * ((T + 255) // 256),), device="cuda", dtype=torch.int32)
    # both variants of the launch: conditioner projection inside the GEMM, and precomputed (what a sampling run issues)
    for name, pcond in (("in-GEMM", None), ("hoisted", den.cond_projection_bf16(cond, plan))):
        ref = None
        for it in range(n):
            skip = torch.empty(B, T, plan.C, device="cuda")
            ops.diffnet_stack_bf16(xb0, e0, e1, skip, cond, plan.w1p_all, plan.woutp_all, table[:, 50:51], plan.bout_all, flags,
                                   plan.dilations, stepbias_batch_stride=0, stepbias_layer_stride=table.stride(0), pcond=pcond)
            if ref is None:
                ref = skip.clone()
            elif not torch.equal(skip, ref):
                bad += 1
                print(f"stack {name} B={B} T={T}: run {it} differs, max|d|={float((skip - ref).abs().max()):.3e}", flush=True)
        torch.cuda.synchronize()
        print(f"stack {name} B={B} T={T}: {n} runs, finite={bool(torch.isfinite(ref).all())}", flush=True)
# uSFGAN block, fixed and adaptive
B, T = 3, 200000
xb = torch.randn(B, T, 64, device="cuda").to(torch.bfloat16); aux = torch.randn(B, T, 80, device="cuda").to(torch.bfloat16)
w1p, woutp = ops.usfgan_pack_block(torch.randn(128, 64, 3, device="cuda") * 0.05, torch.randn(128, 80, device="cuda") * 0.05,
                                   torch.randn(64, 64, device="cuda") * 0.1)
b1 = torch.randn(128, device="cuda") * 0.1; bo = torch.randn(64, device="cuda") * 0.1
idx = ops.pd_index(torch.empty(B, 1, T, device="cuda").uniform_(2, 40), 4)
for name, kw in (("fixed", dict(dilation=8)), ("adaptive", dict(idx=idx))):
    ref = None
    for it in range(n):
        out = torch.empty_like(xb)
        ops.usfgan_block_bf16(xb, out, aux, w1p, woutp, b1, bo, **kw)
        if ref is None:
            ref = out.clone()
        elif not torch.equal(out, ref):
            bad += 1
            print(f"usfgan {name}: run {it} differs", flush=True)
    print(f"usfgan {name}: {n} runs", flush=True)
# LSTM recurrence (cross-CTA st.async hand-off every step), ragged lengths, many clusters in flight
for H in (64, 128, 256, 160):
    B, T = 9, 700
    g = torch.Generator().manual_seed(H)
    pre = torch.randn(B, T, 8 * H, generator=g).cuda()
    w_hh = (torch.randn(2, 4 * H, H, generator=g) / H ** 0.5).cuda()
    lens = torch.tensor([700, 699, 512, 300, 257, 256, 100, 3, 1], dtype=torch.int32, device="cuda")
    ref = None
    for it in range(n):
        h = torch.full((B, 2 * H, T), float("nan"), device="cuda")
        ops.lstm_f32(pre, w_hh, lens, H, pre_layout="ntc", h_f32=h)
        if ref is None:
            ref = h.clone()
        elif not torch.equal(h, ref):
            bad += 1
            print(f"lstm H={H}: run {it} differs", flush=True)
    print(f"lstm H={H}: {n} runs, finite={bool(torch.isfinite(ref).all())}", flush=True)
print("mismatching runs:", bad)
