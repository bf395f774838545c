"""Which start offsets along the innermost (time) dimension does the wgrad kernel's TMA box accept? (debug aid)"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from ensemble_svs_with_interactions_b200 import ops
    s = int(sys.argv[1])
    B, T, N, K = 2, 200, 128, 64
    p = torch.randn(B, T, N, device="cuda").to(torch.bfloat16); q = torch.randn(B, T, K, device="cuda").to(torch.bfloat16)
    dW = torch.empty(N, K, device="cuda")
    ops.wgrad_bf16(ops.ntc_to_nct_bf16(p), [(ops.ntc_to_nct_bf16(q), s)], dW, T=T)
    torch.cuda.synchronize()
    qs = torch.zeros_like(q)
    if s >= 0: qs[:, :T - s] = q[:, s:]
    else: qs[:, -s:] = q[:, :T + s]
    ref = torch.einsum("btn,btk->nk", p.float(), qs.float())
    print(f"shift {s:+d}: max_abs {(dW - ref).abs().max().item():.3e} (ref max {ref.abs().max().item():.2e})")
else:
    for s in (0, 8, -8, 4, -4, 1, -1):
        r = subprocess.run([sys.executable, __file__, str(s)], capture_output=True, text=True)
        out = [l for l in (r.stdout + r.stderr).splitlines() if "shift" in l or "Error" in l or "error" in l]
        print(f"[{s:+d}]", out[-1] if out else "no output", flush=True)
