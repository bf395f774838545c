"""C = 128 residual stack: the two-tiles-per-CTA-pair kernel (diffnet_stack_duo_sm100.cu) against the one-tile kernel
(SVSK_STACK_NO_DUO=1) on the same inputs — the skip sums must be BIT-IDENTICAL (same MMAs in the same order, same
epilogue arithmetic) — with timings of both and a determinism soak.  usage: python tools/check_stack_duo.py [repeats=50]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ensemble_svs_with_interactions_b200 import ops  # noqa: E402
from ensemble_svs_with_interactions_b200.diffsinger import DiffNet, GaussianDiffusion  # noqa: E402
from ensemble_svs_with_interactions_b200.diffsinger import denoiser as _den  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
bad = 0
for (M_, H_, L_, C_) in ((5, 128, 10, 128), (5, 64, 3, 128)):
    torch.manual_seed(0)
    m = GaussianDiffusion(H_, M_, DiffNet(M_, H_, L_, C_, 4), K_step=100).to("cuda").eval()
    plan = m.denoise_fn.bf16_plan()
    table = m._step_table()
    for B, T in ((6, 6000), (6, 2000), (3, 257), (2, 700), (5, 517), (1, 100), (2, 2049), (4, 512)):
        g = torch.Generator().manual_seed(B * T + L_)
        cond = torch.randn(B, T, plan.H, generator=g).cuda().to(torch.bfloat16)
        xb0 = torch.randn(B, T, plan.C, generator=g).cuda().to(torch.bfloat16)
        flags = torch.empty((B * 2 * ((T + 255) // 256),), device="cuda", dtype=torch.int32)

        def run(duo, n=1, timed=False):
            os.environ.pop("SVSK_STACK_NO_DUO", None)
            os.environ.pop("SVSK_STACK_DUO", None)
            os.environ["SVSK_STACK_DUO" if duo else "SVSK_STACK_NO_DUO"] = "1"
            _den._STACK_FIT_CACHE.clear()
            nb = B
            while nb > 0 and not ops.diffnet_stack_fits(nb, T, plan.C, plan.H):
                nb -= 1
            assert nb > 0
            skip = torch.empty(B, T, plan.C, device="cuda")
            e0, e1 = torch.empty_like(xb0), torch.empty_like(xb0)

            def once():
                for b0 in range(0, B, nb):
                    b1 = min(B, b0 + nb)
                    ops.diffnet_stack_bf16(xb0[b0:b1], e0[b0:b1], e1[b0:b1], skip[b0:b1], cond[b0:b1], plan.w1p_all,
                                           plan.woutp_all, table[:, 50:51], plan.bout_all, flags, plan.dilations,
                                           stepbias_batch_stride=0, stepbias_layer_stride=table.stride(0))
            once()
            torch.cuda.synchronize()
            us = None
            if timed:
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n):
                    once()
                b_.record(); b_.synchronize()
                us = a.elapsed_time(b_) * 1e3 / n
            os.environ.pop("SVSK_STACK_NO_DUO", None)
            os.environ.pop("SVSK_STACK_DUO", None)
            return skip, us, nb

        ref, us_ref, nb_ref = run(False, 10, True)
        out, us_duo, nb_duo = run(True, 10, True)
        same = torch.equal(out, ref)
        flops = 2.0 * B * T * (2 * plan.C * (3 * plan.C + plan.H) + 2 * plan.C * plan.C) * plan.L
        print(f"H={H_} L={L_} B={B} T={T}: one tile per pair {us_ref:7.1f} us ({nb_ref} tracks/launch) | two tiles {us_duo:7.1f} us "
              f"({nb_duo} tracks/launch, {flops / us_duo / 1e6:6.1f} TFLOP/s) | bit-identical={same} finite={bool(torch.isfinite(out).all())}",
              flush=True)
        if not same:
            bad += 1
            print(f"   max|d| = {float((out - ref).abs().max()):.3e} of {float(ref.abs().max()):.3e}", flush=True)
        for it in range(reps if (B, T) in ((6, 6000), (5, 517)) else 3):
            o2, _, _ = run(True)
            if not torch.equal(o2, out):
                bad += 1
                print(f"   run {it} differs", flush=True)
                break
print("CHECK_STACK_DUO", "OK" if bad == 0 else f"FAILED ({bad})")
sys.exit(1 if bad else 0)
