"""Unit checks of the training kernels (svsk_seggemm_bf16 modes, svsk_wgrad_bf16, helpers) one launch at a time, each
followed by a synchronize and compared with plain torch on the same bf16-rounded operands (debug aid)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ensemble_svs_with_interactions_b200 import ops  # noqa: E402

dev = "cuda"
bf = torch.bfloat16
torch.manual_seed(0)


def shifted(x, s):
    """x [B,T,K] -> x[:, t + s] with zeros outside."""
    B, T, K = x.shape
    out = torch.zeros_like(x)
    if s >= 0:
        out[:, :T - s] = x[:, s:]
    else:
        out[:, -s:] = x[:, :T + s]
    return out


def report(name, got, ref):
    torch.cuda.synchronize()
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    print(f"{name:34s} max_abs={err:.3e} (ref max {scale:.3e})  {'OK' if err <= 2e-2 * max(scale, 1e-3) else 'MISMATCH'}", flush=True)


def main():
    B, T, C, H = 2, 200, 128, 64
    C2 = 2 * C
    x = torch.randn(B, T, C, device=dev).to(bf)
    w = (torch.randn(96, C, device=dev) / math.sqrt(C)).to(bf)
    bias = torch.randn(96, device=dev)
    # 1. PLAIN
    out = torch.empty(B, T, 96, device=dev, dtype=bf)
    outf = torch.empty(B, T, 96, device=dev)
    ops.seggemm_bf16([(x, C, 0)], w, mode=ops.SEG_PLAIN, bias=bias, act=ops.ACT_RELU, out0=out, outf=outf)
    ref = torch.relu(x.float() @ w.float().t() + bias)
    report("seggemm PLAIN relu (bf16 out)", out, ref)
    report("seggemm PLAIN relu (fp32 out)", outf, ref)
    # 2. three shifted taps + second tensor
    c = torch.randn(B, T, H, device=dev).to(bf)
    w3 = (torch.randn(C2, 3 * C + H, device=dev) / math.sqrt(3 * C + H)).to(bf)
    d = 4
    xin = torch.cat([shifted(x, -d), x, shifted(x, d), c], dim=-1).float()
    b1 = torch.randn(C2, device=dev) * 0.1
    ypre = torch.empty(B, T, C2, device=dev, dtype=bf)
    z = torch.empty(B, T, C, device=dev, dtype=bf)
    ops.seggemm_bf16([(x, C, -d), (x, C, 0), (x, C, d), (c, H, 0)], w3, mode=ops.SEG_GATE_FWD, bias=b1, out0=ypre, out1=z)
    yref = xin @ w3.float().t() + b1
    report("seggemm GATE_FWD ypre", ypre, yref)
    zref = torch.cat([torch.sigmoid(yref[..., p * 256:p * 256 + 128]) * torch.tanh(yref[..., p * 256 + 128:p * 256 + 256])
                      for p in range(C2 // 256)], dim=-1)
    report("seggemm GATE_FWD z", z, zref)
    # 3. RES_SKIP
    wo = (torch.randn(C2, C, device=dev) / math.sqrt(C)).to(bf)
    bo = torch.randn(C2, device=dev) * 0.1
    dpn = torch.randn(B, C, device=dev)
    xn = torch.empty(B, T, C, device=dev, dtype=bf); xdn = torch.empty_like(xn)
    skip = torch.ones(B, T, C, device=dev)
    ops.seggemm_bf16([(z, C, 0)], wo, mode=ops.SEG_RES_SKIP, bias=bo, in0=x, out0=xn, out1=xdn, dp_next=dpn, outf=skip, init=False)
    o = z.float() @ wo.float().t() + bo
    report("seggemm RES_SKIP x'", xn, (x.float() + o[..., :C]) / math.sqrt(2))
    report("seggemm RES_SKIP xd'", xdn, (x.float() + o[..., :C]) / math.sqrt(2) + dpn[:, None])
    report("seggemm RES_SKIP skip +=", skip, 1.0 + o[..., C:])
    # 4. GATE_BWD
    u = torch.randn(B, T, C, device=dev).to(bf); dS = torch.randn(B, T, C, device=dev).to(bf)
    woT = wo.t().contiguous()
    dy = torch.empty(B, T, C2, device=dev, dtype=bf)
    ops.seggemm_bf16([(u, C, 0), (dS, C, 0)], woT, mode=ops.SEG_GATE_BWD, in0=ypre, out0=dy)
    dz = torch.cat([u, dS], -1).float() @ wo.float()
    yp = ypre.float()
    dyref = torch.empty(B, T, C2, device=dev)
    for p in range(C2 // 256):
        g, f = yp[..., p * 256:p * 256 + 128], yp[..., p * 256 + 128:p * 256 + 256]
        s, th = torch.sigmoid(g), torch.tanh(f)
        dzp = dz[..., p * 128:(p + 1) * 128]
        dyref[..., p * 256:p * 256 + 128] = dzp * th * s * (1 - s)
        dyref[..., p * 256 + 128:p * 256 + 256] = dzp * s * (1 - th * th)
    report("seggemm GATE_BWD dy", dy, dyref)
    # 5. ADD_SCALE with mask
    w1T = (torch.randn(C, 3 * C2, device=dev) / math.sqrt(3 * C2)).to(bf)
    un = torch.empty(B, T, C, device=dev, dtype=bf)
    ops.seggemm_bf16([(dy, C2, d), (dy, C2, 0), (dy, C2, -d)], w1T, mode=ops.SEG_ADD_SCALE, in0=u, alpha=0.5, mask=x, out0=un)
    acc = torch.cat([shifted(dy, d), dy, shifted(dy, -d)], -1).float() @ w1T.float().t()
    report("seggemm ADD_SCALE masked", un, torch.where(x.float() > 0, (acc + u.float()) * 0.5, torch.zeros_like(acc)))
    # 6. transposes + wgrad (two track groups; indicator rows give the column sums)
    dyT = ops.ntc_to_nct_bf16(dy); cT = ops.ntc_to_nct_bf16(c)
    report("ntc_to_nct", dyT[:, :, :T], dy.transpose(1, 2))
    xT3 = ops.ntc_to_nct_bf16(x, shifts=(-d, 0, d))
    report("ntc_to_nct shifted -d", xT3[:, :C, :T], shifted(x, -d).transpose(1, 2))
    report("ntc_to_nct shifted +d", xT3[:, 2 * C:, :T], shifted(x, d).transpose(1, 2))
    Tp = dyT.shape[2]
    ind = torch.zeros(B, 16, Tp, device=dev, dtype=bf)
    for b in range(B):
        ind[b, 3 * b, :T] = 1; ind[b, 3 * b + 1, :d] = 1; ind[b, 3 * b + 2, T - d:T] = 1
    dW = torch.empty(2, C2, 3 * C + H + 16, device=dev)
    ops.wgrad_bf16(dyT, [(xT3, 0), (cT, 0), (ind, 0)], dW, T=T)
    dWs = dW.sum(0)
    report("wgrad (3 taps + cond), 2 groups", dWs[:, :3 * C + H], torch.einsum("btn,btk->nk", dy.float(), xin))
    cs = dWs[:, 3 * C + H:3 * C + H + 3 * B].reshape(C2, B, 3)
    report("indicator rows: sum all", cs[:, :, 0].t(), dy.float().sum(1))
    report("indicator rows: sum head", cs[:, :, 1].t(), dy[:, :d].float().sum(1))
    report("indicator rows: sum tail", cs[:, :, 2].t(), dy[:, T - d:].float().sum(1))


if __name__ == "__main__":
    main()
