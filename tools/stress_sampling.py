"""Race hunt over whole sampling passes (debug aid): GaussianDiffusion.inference with injected x_T / z, CUDA-graph replay and
eager launches, at BASELINE config 2 and a long-track shape — every pass must be bit-identical to the first."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
m = bench.build_model().to("cuda").eval()
bad = 0
for B, T in ((6, 2000), (2, 6000), (3, 517)):
    g = torch.Generator().manual_seed(B * T)
    cond = torch.randn(B, T, 256, generator=g).cuda()
    x_T = torch.randn(B, 1, 80, T, generator=g).cuda()
    z = torch.randn(m.K_step, B, 1, 80, T, generator=g).cuda()
    ref = None
    for graph in (True, False):
        m.use_cuda_graph = graph
        for it in range(n if graph else 2):
            y = m.inference(cond, x_T=x_T, z=z)
            if ref is None:
                ref = y.clone()
            elif not torch.equal(y, ref):
                bad += 1
                print(f"B={B} T={T} graph={graph}: pass {it} differs, max|d|={float((y - ref).abs().max()):.3e}", flush=True)
    torch.cuda.synchronize()
    print(f"sampling B={B} T={T}: {n} graph replays + 2 eager passes, finite={bool(torch.isfinite(ref).all())}", flush=True)
print("mismatching passes:", bad)
print("STRESS_SAMPLING OK" if bad == 0 else "STRESS_SAMPLING FAILED")
