"""Per-CTA clock64 timeline of the fused DiffNet block kernel (profiling aid; not part of the product path)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ensemble_svs_with_interactions_b200 import ops  # noqa: E402

B, T = bench.B, bench.T
m = bench.build_model().to("cuda")
plan = m.denoise_fn.bf16_plan()
cond = torch.randn(B, T, 256, device="cuda").to(torch.bfloat16)
sb = [tl[50] for tl in m._step_table()]
xb0 = torch.randn(B, T, plan.C, device="cuda").to(torch.bfloat16)
xb1 = torch.empty_like(xb0)
x32 = torch.randn(B, T, plan.C, device="cuda")
skip32 = torch.zeros(B, T, plan.C, device="cuda")
names = ["start", "loads issued", "G1 p0 issued", "G1 p1 issued", "G ready@mma", "G2 p0 issued", "G2 p1 issued",
         "D1 p0 full", "gate p0 done", "D1 p1 full", "gate p1 done", "gating done", "D2 p0 full", "D2 p1 full",
         "epi done", "end"]


def run(layer, tile, ablate, kernel=1):
    lw = plan.layers[layer]
    os.environ["SVSK_DIFFNET_ABLATE"] = str(ablate)
    os.environ.pop("SVSK_DIFFNET_TIMELINE", None)
    def launch():
        ops.diffnet_block_bf16(xb0, xb1, x32, skip32, cond, lw["w1p"], lw["woutp"], sb[layer], lw["bout"],
                               dilation=lw["dilation"], stepbias_batch_stride=0, init_skip=False, write_x=True,
                               time_tile=tile, kernel=kernel)
    for _ in range(5):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        launch()
    e1.record(); e1.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 50
    dbg = torch.zeros(512 * 32, dtype=torch.int64, device="cuda")
    os.environ["SVSK_DIFFNET_TIMELINE"] = str(dbg.data_ptr())
    launch()
    torch.cuda.synchronize()
    os.environ.pop("SVSK_DIFFNET_TIMELINE", None)
    d = dbg.view(512, 32).cpu()
    d = d[d[:, 15] > 0]
    if kernel >= 2:
        d = d[d[:, 2] > 0]   # leader CTAs carry the MMA stamps
    rel = (d[:, :16] - d[:, :1]).float()
    med = rel.median(dim=0).values
    acc = d[:, 16:20].float().median(dim=0).values
    print(f"--- kernel {kernel} layer {layer} (dil {lw['dilation']}) tile {tile} ablate {ablate}: {us:.2f} us/launch, {d.shape[0]} CTAs; "
          f"median cycles since CTA start:")
    print("   " + "  ".join(f"{n}={int(v)}" for n, v in zip(names, med)))
    if kernel == 3:
        st = (d[:, 19:24] - d[:, :1]).float().median(dim=0).values
        print("   kernel 3 (leader CTA, cycles): producer wait empty=%d | mma blocking waits (probe misses)=%d | "
              "grid dependency resolved @%d | MMA thread starts @%d | first cond tile landed @%d | cond k-blocks issued @%d | window landed @%d"
              % (int(acc[0]), int(acc[1]), int(st[2]), int(st[3]), int(st[4]), int(st[0]), int(st[1])))
    if kernel == 2:
        print("   GEMM1 accounting (leader CTA, cycles): producer wait empty=%d | mma wait own stage=%d | mma wait peer stage=%d | "
              "mma issue+commit=%d" % tuple(int(v) for v in acc))


for ab in (0,):
    run(1, 0, ab, kernel=3)
os.environ['SVSK_NO_PDL'] = '1'
run(1, 0, 0, kernel=3)
os.environ.pop('SVSK_NO_PDL')
run(1, 0, 0, kernel=2)
