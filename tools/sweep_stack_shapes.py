"""Robustness sweep (debug aid): one-launch residual stack + fused step kernel vs the layer-at-a-time kernels on odd shapes."""
import os, sys, itertools, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ensemble_svs_with_interactions_b200.diffsinger import DiffNet, GaussianDiffusion

def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-12))

torch.manual_seed(0)
bad = 0
for (C, H, M, L), (B, T) in itertools.product([(256, 256, 80, 3), (128, 192, 60, 5), (128, 64, 5, 2), (256, 64, 33, 4)],
                                              [(1, 1), (2, 7), (1, 8), (3, 9), (2, 127), (1, 128), (2, 129), (1, 255), (2, 256), (1, 257), (2, 2049), (1, 4100)]):
    den = DiffNet(M, H, L, C, 4)
    with torch.no_grad():
        den.output_projection.weight.normal_(0, 0.05)
    m = GaussianDiffusion(H, M, den, K_step=3).to("cuda").eval()
    m.use_cuda_graph = False
    cond = torch.randn(B, T, H, device="cuda"); x_T = torch.randn(B, 1, M, T, device="cuda"); z = torch.randn(3, B, 1, M, T, device="cuda")
    try:
        y = m.inference(cond, x_T=x_T, z=z)
        os.environ["SVSK_DIFFNET_STACK"] = "0"; os.environ["SVSK_DIFFNET_STEP"] = "0"
        ref = m.inference(cond, x_T=x_T, z=z)
        os.environ.pop("SVSK_DIFFNET_STACK"); os.environ.pop("SVSK_DIFFNET_STEP")
        torch.cuda.synchronize()
        r = rel(y, ref)
        ok = bool(torch.isfinite(y).all()) and r < 2e-2
    except Exception as e:  # noqa: BLE001
        ok, r = False, str(e)[:120]
        os.environ.pop("SVSK_DIFFNET_STACK", None); os.environ.pop("SVSK_DIFFNET_STEP", None)
    bad += not ok
    print(f"C={C} H={H} M={M} L={L} B={B} T={T}: {'ok' if ok else 'FAIL'} {r}", flush=True)
print("failures:", bad)
