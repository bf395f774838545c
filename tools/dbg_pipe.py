"""Debug aid: small DiffNet shapes through GaussianDiffusion.inference (eager and CUDA graph)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
from ensemble_svs_with_interactions_b200.diffsinger import DiffNet, GaussianDiffusion
torch.manual_seed(11)
models = []
for name, args, H, M in (("mgc", (60, 128, 4, 128, 4), 128, 60), ("bap", (5, 64, 2, 128, 2), 64, 5)):
    m = GaussianDiffusion(H, M, DiffNet(*args), K_step=6).to("cuda").eval()
    with torch.no_grad():
        m.denoise_fn.output_projection.weight.normal_(0, 0.05)
    models.append((name, m, H))
for graph in (False, True):
    for B, T in ((3, 64), (2, 64), (1, 64), (3, 52), (2, 40), (1, 17), (3, 64)):
        for name, m, H in models:
            m.use_cuda_graph = graph
            try:
                y = m.inference(torch.randn(B, T, H, device="cuda"))
                torch.cuda.synchronize()
                print(name, "graph", graph, B, T, "ok", float(y.abs().mean()), flush=True)
            except Exception as e:
                print(name, "graph", graph, B, T, "FAILED", str(e)[:300], flush=True)
                sys.exit(1)
