"""TEST INFRASTRUCTURE ONLY — CPU oracle for the gated dilated-conv hot path.

This file is a *restatement* of the reference's algorithm (sarulab-speech/
ensemble_svs_with_interactions, pure Python/PyTorch) written from its behaviour,
function by function, as stateless fp32 functions over plain ``state_dict``
tensors.  It is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product package never imports ``oracle``.

Why torch and not numpy/C: the path is floating-point convolution arithmetic
(no byte/integer work except gather indices); ``torch.nn.functional`` on CPU is
the same fp32 arithmetic (oneDNN) the reference itself runs on a CPU host, which
is what the CPU baseline has to time.  Only *functional* torch ops are used.

Parity is PINNED: ``tests/test_oracle_golden.py`` checks every function below
against ``tests/golden/*.npz`` — outputs of the unmodified reference modules,
generated in the build container by ``oracle/make_golden.py`` (the reference's
own tests hold no golden vectors for this path, SURVEY.md §4/§8c: shape checks
only), and ``tests/test_oracle_vs_reference.py`` re-checks against the live
reference whenever ``/root/reference`` is present.

Each function cites the reference ``file:line`` it follows (paths relative to
``/root/reference``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------- #
# helpers
# --------------------------------------------------------------------------- #
def fold_weight_norm(sd: SD, prefix: str) -> Tensor:
    """Effective conv weight for ``prefix`` whether or not weight-norm is applied.

    Old-style ``torch.nn.utils.weight_norm`` keeps ``weight_g`` (Cout,1,..) and
    ``weight_v``; w = g * v / ||v|| with the norm over every dim but 0
    (nnsvs/wavenet/modules.py:6-9,
    nnsvs/usfgan/models/generator.py:158-166).
    """
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"]
    g = sd[prefix + ".weight_g"]
    v = sd[prefix + ".weight_v"]
    dims = tuple(range(1, v.dim()))
    return v * (g / v.norm(2, dim=dims, keepdim=True))


def _bias(sd: SD, prefix: str) -> Optional[Tensor]:
    return sd.get(prefix + ".bias", None)


def shifted_tap(x: Tensor, shift: int, mode: str) -> Tensor:
    """x[..., t + shift] for every t, with out-of-range indices resolved by ``mode``.

    mode: "zeros" (0 outside), "reflect" (i<0 -> -i, i>=T -> 2(T-1)-i; no edge
    repeat), "replicate" (clamp).
    """
    T = x.shape[-1]
    idx = torch.arange(T) + shift
    if mode == "zeros":
        valid = (idx >= 0) & (idx < T)
        out = x[..., idx.clamp(0, T - 1)]
        return out * valid.to(x.dtype)
    if mode == "reflect":
        idx = torch.where(idx < 0, -idx, idx)
        idx = torch.where(idx >= T, 2 * (T - 1) - idx, idx)
        return x[..., idx]
    if mode == "replicate":
        return x[..., idx.clamp(0, T - 1)]
    raise ValueError(mode)


def conv_taps(x: Tensor, w: Tensor, b: Optional[Tensor], shifts: Sequence[int], mode: str) -> Tensor:
    """Conv1d as a sum of per-tap 1x1 products: y = b + sum_j w[:,:,j] @ x[t+shifts[j]]."""
    y = None
    for j, s in enumerate(shifts):
        term = torch.einsum("oc,bct->bot", w[:, :, j], shifted_tap(x, s, mode))
        y = term if y is None else y + term
    if b is not None:
        y = y + b[None, :, None]
    return y


def conv1x1(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    y = torch.einsum("oc,bct->bot", w[:, :, 0], x)
    if b is not None:
        y = y + b[None, :, None]
    return y


# --------------------------------------------------------------------------- #
# DiffNet denoiser  (nnsvs/diffsinger/denoiser.py)
# --------------------------------------------------------------------------- #
def mish(x: Tensor) -> Tensor:
    """denoiser.py:9-11."""
    return x * torch.tanh(F.softplus(x))


def sinusoidal_embedding(t: Tensor, dim: int) -> Tensor:
    """denoiser.py:14-26: cat(sin(t*f), cos(t*f)), f_i = 10000^(-i/(dim/2-1))."""
    half = dim // 2
    freq = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    arg = t[:, None] * freq[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def diffnet_step_embedding(sd: SD, t: Tensor, channels: int) -> Tensor:
    """denoiser.py:113-114: mlp(SinusoidalPosEmb(t)) -> (B, C)."""
    e = sinusoidal_embedding(t, channels)
    e = F.linear(e, sd["mlp.0.weight"], sd["mlp.0.bias"])
    e = mish(e)
    return F.linear(e, sd["mlp.2.weight"], sd["mlp.2.bias"])


def diffnet_block(sd: SD, i: int, x: Tensor, cond: Tensor, emb: Tensor, dilation: int) -> Tuple[Tensor, Tensor]:
    """One gated block, denoiser.py:54-66.

    y = conv3_dil_zero_pad(x + proj(e)) + cond1x1(c); gate,filt = halves(y);
    z = sigmoid(gate) * tanh(filt); r,s = halves(out1x1(z)); ((x+r)/sqrt2, s).
    NB: sigmoid on the FIRST half (opposite to WaveNet/uSFGAN).
    """
    p = f"residual_layers.{i}."
    step = F.linear(emb, sd[p + "diffusion_projection.weight"], sd[p + "diffusion_projection.bias"])
    y = x + step[:, :, None]
    y = conv_taps(y, sd[p + "dilated_conv.weight"], sd[p + "dilated_conv.bias"],
                  (-dilation, 0, dilation), "zeros")
    y = y + conv1x1(cond, sd[p + "conditioner_projection.weight"], sd[p + "conditioner_projection.bias"])
    C = x.shape[1]
    z = torch.sigmoid(y[:, :C]) * torch.tanh(y[:, C:])
    o = conv1x1(z, sd[p + "output_projection.weight"], sd[p + "output_projection.bias"])
    return (x + o[:, :C]) / math.sqrt(2.0), o[:, C:]


def diffnet_forward(sd: SD, spec: Tensor, t: Tensor, cond: Tensor,
                    residual_layers: int, dilation_cycle_length: int) -> Tensor:
    """DiffNet.forward, denoiser.py:101-124.  spec (B,1,M,T), t (B,), cond (B,H,T)."""
    C = sd["input_projection.weight"].shape[0]
    x = F.relu(conv1x1(spec[:, 0], sd["input_projection.weight"], sd["input_projection.bias"]))
    emb = diffnet_step_embedding(sd, t.to(torch.float32), C)
    skip_sum = torch.zeros_like(x)
    for i in range(residual_layers):
        x, s = diffnet_block(sd, i, x, cond, emb, 2 ** (i % dilation_cycle_length))
        skip_sum = skip_sum + s
    x = skip_sum / math.sqrt(residual_layers)
    x = F.relu(conv1x1(x, sd["skip_projection.weight"], sd["skip_projection.bias"]))
    x = conv1x1(x, sd["output_projection.weight"], sd["output_projection.bias"])
    return x[:, None]


# --------------------------------------------------------------------------- #
# Gaussian diffusion  (nnsvs/diffsinger/diffusion.py)
# --------------------------------------------------------------------------- #
SCHEDULE_BUFFERS = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod",
    "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
    "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2",
)


def beta_schedule(K: int, schedule_type: str = "linear", **params) -> np.ndarray:
    """diffusion.py:27-51 (float64)."""
    if schedule_type == "linear":
        return np.linspace(params.get("min_beta", 1e-4), params.get("max_beta", 0.06), K)
    if schedule_type == "cosine":
        s = params.get("s", 0.008)
        n = K + 1
        x = np.linspace(0, n, n)
        ac = np.cos(((x / n) + s) / (1 + s) * np.pi * 0.5) ** 2
        ac = ac / ac[0]
        return np.clip(1 - ac[1:] / ac[:-1], a_min=0, a_max=0.999)
    raise ValueError(schedule_type)


def diffusion_tables(betas: np.ndarray) -> Dict[str, Tensor]:
    """The 12 fp32 schedule buffers, computed in float64 (diffusion.py:98-145)."""
    betas = np.asarray(betas, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas)
    ac_prev = np.append(1.0, ac[:-1])
    pv = betas * (1.0 - ac_prev) / (1.0 - ac)
    tab = {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": np.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1),
        "posterior_variance": pv,
        "posterior_log_variance_clipped": np.log(np.maximum(pv, 1e-20)),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }
    return {k: torch.tensor(v, dtype=torch.float32) for k, v in tab.items()}


def _at(table: Tensor, t: Tensor, ndim: int) -> Tensor:
    """diffusion.py:10-13 ``extract``."""
    return table[t].reshape(t.shape[0], *([1] * (ndim - 1)))


def ddpm_update(tab: Dict[str, Tensor], x: Tensor, t: Tensor, eps: Tensor, z: Tensor,
                clip_denoised: bool = True) -> Tensor:
    """One ancestral step given the denoiser output (diffusion.py:164-204).

    x0 = a_t x - b_t eps ; clamp(+-1) ; mean = c1_t x0 + c2_t x ;
    x <- mean + [t>0] exp(0.5 logvar_t) z
    """
    n = x.dim()
    x0 = _at(tab["sqrt_recip_alphas_cumprod"], t, n) * x - _at(tab["sqrt_recipm1_alphas_cumprod"], t, n) * eps
    if clip_denoised:
        x0 = x0.clamp(-1.0, 1.0)
    mean = _at(tab["posterior_mean_coef1"], t, n) * x0 + _at(tab["posterior_mean_coef2"], t, n) * x
    logvar = _at(tab["posterior_log_variance_clipped"], t, n)
    nonzero = (1 - (t == 0).float()).reshape(t.shape[0], *([1] * (n - 1)))
    return mean + nonzero * (0.5 * logvar).exp() * z


def q_sample(tab: Dict[str, Tensor], x0: Tensor, t: Tensor, noise: Tensor) -> Tensor:
    """diffusion.py:261-267."""
    n = x0.dim()
    return _at(tab["sqrt_alphas_cumprod"], t, n) * x0 + _at(tab["sqrt_one_minus_alphas_cumprod"], t, n) * noise


def diffusion_inference(sd: SD, cond: Tensor, x_T: Tensor, z: Tensor, *, K_step: int,
                        residual_layers: int, dilation_cycle_length: int, norm_scale: float = 10.0,
                        return_trajectory: bool = False, keep_steps: Optional[Sequence[int]] = None):
    """GaussianDiffusion.inference with INJECTED noise (diffusion.py:302-336, encoder=None).

    ``sd`` is the GaussianDiffusion state_dict (12 buffers + ``denoise_fn.*``).
    cond (B,T,H); x_T (B,1,M,T); z (K,B,1,M,T) with z[i] consumed at step t=i.
    Returns (B,T,M) [, list of per-step (eps, x)]; with ``keep_steps`` the second value is {t: (eps, x)} for those
    steps only (full-size runs: 100 x 2 tensors would not fit comfortably).
    """
    den = {k[len("denoise_fn."):]: v for k, v in sd.items() if k.startswith("denoise_fn.")}
    tab = {k: sd[k] for k in SCHEDULE_BUFFERS}
    c = cond.transpose(1, 2)
    x = x_T
    B = x.shape[0]
    traj = []
    for i in reversed(range(K_step)):
        t = torch.full((B,), i, dtype=torch.long)
        eps = diffnet_forward(den, x, t, c, residual_layers, dilation_cycle_length)
        x = ddpm_update(tab, x, t, eps, z[i])
        if keep_steps is not None:
            if i in keep_steps:
                traj.append((i, (eps, x)))
        elif return_trajectory:
            traj.append((eps, x))
    out = x[:, 0].transpose(1, 2) * norm_scale
    if keep_steps is not None:
        return out, dict(traj)
    return (out, traj) if return_trajectory else out


def diffusion_training_forward(sd: SD, cond: Tensor, y: Tensor, t: Tensor, noise: Tensor, *,
                               residual_layers: int, dilation_cycle_length: int,
                               norm_scale: float = 10.0) -> Tuple[Tensor, Tensor]:
    """GaussianDiffusion.forward with INJECTED t and noise (diffusion.py:269-300).

    cond (B,T,H), y (B,T,M), t (B,), noise (B,1,M,T) -> (noise, eps_hat) as (B,T,M).
    """
    den = {k[len("denoise_fn."):]: v for k, v in sd.items() if k.startswith("denoise_fn.")}
    tab = {k: sd[k] for k in SCHEDULE_BUFFERS}
    x = (y / norm_scale).transpose(1, 2)[:, None]
    x_noisy = q_sample(tab, x, t, noise)
    eps = diffnet_forward(den, x_noisy, t, cond.transpose(1, 2), residual_layers, dilation_cycle_length)
    return noise.squeeze(1).transpose(1, 2), eps.squeeze(1).transpose(1, 2)


def plms_x_pred(tab: Dict[str, Tensor], x: Tensor, noise_t: Tensor, t: Tensor, interval: int) -> Tensor:
    """PLMS/DDIM transfer ``get_x_pred`` (diffusion.py:213-230)."""
    n = x.dim()
    a_t = _at(tab["alphas_cumprod"], t, n)
    a_prev = _at(tab["alphas_cumprod"], torch.clamp(t - interval, min=0), n)
    a_t_sq, a_prev_sq = a_t.sqrt(), a_prev.sqrt()
    delta = (a_prev - a_t) * (
        (1 / (a_t_sq * (a_t_sq + a_prev_sq))) * x
        - 1 / (a_t_sq * (((1 - a_prev) * a_t).sqrt() + ((1 - a_t) * a_prev).sqrt())) * noise_t
    )
    return x + delta


# --------------------------------------------------------------------------- #
# WaveNet  (nnsvs/wavenet/modules.py, wavenet.py)
# --------------------------------------------------------------------------- #
def wavenet_block(sd: SD, i: int, x: Tensor, c: Tensor, dilation: int) -> Tuple[Tensor, Tensor]:
    """ResSkipBlock._forward (non-incremental), modules.py:88-122.

    Causal: pad (k-1)d both sides then trim right == taps at t-(k-1-j)d with zeros
    for negative indices.  tanh on FIRST half, sigmoid on second.  x + res with no
    sqrt(1/2) scale.
    """
    p = f"main_conv_layers.{i}."
    w = fold_weight_norm(sd, p + "conv")
    k = w.shape[2]
    shifts = [-(k - 1 - j) * dilation for j in range(k)]
    y = conv_taps(x, w, _bias(sd, p + "conv"), shifts, "zeros")
    y = y + conv1x1(c, fold_weight_norm(sd, p + "conv1x1c"), None)
    h = y.shape[1] // 2
    z = torch.tanh(y[:, :h]) * torch.sigmoid(y[:, h:])
    s = conv1x1(z, fold_weight_norm(sd, p + "conv1x1_skip"), _bias(sd, p + "conv1x1_skip"))
    o = conv1x1(z, fold_weight_norm(sd, p + "conv1x1_out"), _bias(sd, p + "conv1x1_out"))
    return o + x, s


def wavenet_forward(sd: SD, c: Tensor, x: Tensor, layers: int, stacks: int = 1) -> Tensor:
    """WaveNet.forward, wavenet.py:60-87.  c (B,T,Cin), x (B,T,Cout) -> (B,T,Cout)."""
    x = x.transpose(1, 2)
    c = c.transpose(1, 2)
    x = conv1x1(x, fold_weight_norm(sd, "first_conv"), _bias(sd, "first_conv"))
    per_stack = layers // stacks
    skips = 0
    for i in range(layers):
        x, s = wavenet_block(sd, i, x, c, 2 ** (i % per_stack))
        skips = skips + s
    x = F.relu(skips)
    x = conv1x1(x, fold_weight_norm(sd, "last_conv_layers.1"), _bias(sd, "last_conv_layers.1"))
    x = F.relu(x)
    x = conv1x1(x, fold_weight_norm(sd, "last_conv_layers.3"), _bias(sd, "last_conv_layers.3"))
    return x.transpose(1, 2)


# --------------------------------------------------------------------------- #
# uSFGAN  (nnsvs/usfgan/layers/residual_block.py, upsample.py, models/generator.py)
# --------------------------------------------------------------------------- #
def pd_gather(x: Tensor, d: Tensor, dilation: int) -> Tuple[Tensor, Tensor]:
    """Pitch-dependent past/future taps (nnsvs/usfgan/utils/index.py:12-54).

    delta_t = round_half_even(d_t * dilation) in fp32; xP[t] = x[t-delta_t] (0 if <0),
    xF[t] = x[t+delta_t] (0 if >=T); same delta for every channel.
    The reference computes round(-d*dil + (t-T)) and round(d*dil + t) on floats; for the
    value ranges here (|t| < 2^23) that equals t -/+ round(d*dil) except on exact .5 ties of
    the SUM, so we restate it on the sum exactly as the reference does.
    """
    B, _, T = d.shape
    dil = d * dilation
    ar = torch.arange(T, dtype=torch.float32)
    idx_p = (-dil + (ar - T)).round().long() + T       # absolute index, may be < 0
    idx_f = (dil + ar).round().long()                  # may be >= T
    okp = idx_p >= 0
    okf = idx_f < T
    xp = torch.gather(x, 2, idx_p.clamp(0, T - 1).expand(-1, x.shape[1], -1)) * okp.to(x.dtype)
    xf = torch.gather(x, 2, idx_f.clamp(0, T - 1).expand(-1, x.shape[1], -1)) * okf.to(x.dtype)
    return xp, xf


def _gated_tail(sd: SD, p: str, y: Tensor, residual: Tensor, c: Optional[Tensor]) -> Tensor:
    """Shared tail of Fixed/AdaptiveBlock (residual_block.py:139-157, 217-234); skip conv is dead
    work (ResidualBlocks.forward returns x only, :333-336) and is not computed."""
    if c is not None:
        y = y + conv1x1(c, fold_weight_norm(sd, p + "conv1x1_aux"), None)
    h = y.shape[1] // 2
    z = torch.tanh(y[:, :h]) * torch.sigmoid(y[:, h:])
    o = conv1x1(z, fold_weight_norm(sd, p + "conv1x1_out"), _bias(sd, p + "conv1x1_out"))
    return (o + residual) * math.sqrt(0.5)


def usfgan_fixed_block(sd: SD, p: str, x: Tensor, c: Tensor, dilation: int) -> Tensor:
    """FixedBlock.forward, residual_block.py:123-157 (k=3 reflect-padded dilated conv)."""
    w = fold_weight_norm(sd, p + "conv")
    k = w.shape[2]
    half = (k - 1) // 2
    shifts = [(j - half) * dilation for j in range(k)]
    y = conv_taps(x, w, _bias(sd, p + "conv"), shifts, "reflect")
    return _gated_tail(sd, p, y, x, c)


def usfgan_adaptive_block(sd: SD, p: str, x: Tensor, c: Tensor, d: Tensor, dilation: int) -> Tensor:
    """AdaptiveBlock.forward + pd_indexing, residual_block.py:198-234, 325-332."""
    xp, xf = pd_gather(x, d, dilation)
    y = (conv1x1(x, fold_weight_norm(sd, p + "convC"), _bias(sd, p + "convC"))
         + conv1x1(xp, fold_weight_norm(sd, p + "convP"), _bias(sd, p + "convP"))
         + conv1x1(xf, fold_weight_norm(sd, p + "convF"), _bias(sd, p + "convF")))
    return _gated_tail(sd, p, y, x, c)


def usfgan_residual_blocks(sd: SD, prefix: str, x: Tensor, c: Tensor, d: Tensor, *, blockA: int, cycleA: int,
                           blockF: int, cycleF: int, cascade_mode: int = 0) -> Tensor:
    """ResidualBlocks.forward, residual_block.py:311-336 (+ ctor :274-309 for the dilation plan)."""
    cycleA = max(cycleA, 1)
    cycleF = max(cycleF, 1)
    a_per = blockA // cycleA
    f_per = blockF // cycleF
    modes = [True] * blockA + [False] * blockF if cascade_mode == 0 else [False] * blockF + [True] * blockA
    ia = 0
    jf = 0
    for n, adaptive in enumerate(modes):
        p = f"{prefix}conv_dilated.{n}."
        if adaptive:
            x = usfgan_adaptive_block(sd, p, x, c, d, 2 ** (ia % a_per))
            ia += 1
        else:
            x = usfgan_fixed_block(sd, p, x, c, 2 ** (jf % f_per))
            jf += 1
    return x


def usfgan_periodicity(sd: SD, prefix: str, c: Tensor, conv_layers: int = 3, kernel_size: int = 5,
                       dilation: int = 1, padding_mode: str = "replicate") -> Tensor:
    """PeriodicityEstimator, residual_block.py:339-399: k=5 convs, ReLU..ReLU, Sigmoid."""
    half = kernel_size // 2
    shifts = [(j - half) * dilation for j in range(kernel_size)]
    mode = {"replicate": "replicate", "reflect": "reflect", "zeros": "zeros"}[padding_mode]
    x = c
    for i in range(conv_layers):
        p = f"{prefix}layers.{2 * i}"
        x = conv_taps(x, fold_weight_norm(sd, p), _bias(sd, p), shifts, mode)
        x = torch.sigmoid(x) if i == conv_layers - 1 else F.relu(x)
    return x


def usfgan_upsample(sd: SD, prefix: str, c: Tensor, upsample_scales: Sequence[int]) -> Tensor:
    """ConvInUpsampleNetwork (non-causal), upsample.py:131-194 + 61-128.

    conv_in: k=2w+1, no padding, no bias; then per scale s: nearest repeat xs, then a
    single-channel (1, 2s+1) smoothing filter with zero padding s, shared over aux channels.
    """
    w = fold_weight_norm(sd, prefix + "conv_in")
    c = F.conv1d(c, w)  # valid conv, frame rate
    for n, s in enumerate(upsample_scales):
        c = torch.repeat_interleave(c, s, dim=-1)
        k = fold_weight_norm(sd, f"{prefix}upsample.up_layers.{2 * n + 1}")  # (1,1,1,2s+1)
        taps = k.reshape(-1)
        acc = torch.zeros_like(c)
        for j in range(2 * s + 1):
            acc = acc + taps[j] * shifted_tap(c, j - s, "zeros")
        c = acc
    return c


def _conv_last(sd: SD, x: Tensor) -> Tensor:
    """conv_last = ReLU, 1x1, ReLU, 1x1 (generator.py:461-466)."""
    x = F.relu(x)
    x = conv1x1(x, fold_weight_norm(sd, "conv_last.1"), _bias(sd, "conv_last.1"))
    x = F.relu(x)
    return conv1x1(x, fold_weight_norm(sd, "conv_last.3"), _bias(sd, "conv_last.3"))


def parallel_hn_usfgan_forward(sd: SD, x: Tensor, c: Tensor, d: Tensor, *, harmonic: dict, noise: dict,
                               filt: dict, upsample_scales: Sequence[int] = (5, 4, 3, 2),
                               pe: Optional[dict] = None):
    """ParallelHnUSFGANGenerator.forward, generator.py:472-522.  Returns (x, s, h, n, a)."""
    pe = pe or {}
    c = usfgan_upsample(sd, "upsample_net.", c, upsample_scales)
    assert c.shape[-1] == x.shape[-1]
    a = usfgan_periodicity(sd, "periodicity_estimator.", c, **pe)
    sine, nz = x[:, :1], x[:, 1:2]
    h = conv1x1(sine, fold_weight_norm(sd, "conv_first_sine"), _bias(sd, "conv_first_sine"))
    n = conv1x1(nz, fold_weight_norm(sd, "conv_first_noise"), _bias(sd, "conv_first_noise"))
    h = usfgan_residual_blocks(sd, "harmonic_network.", h, c, d, **harmonic)
    n = usfgan_residual_blocks(sd, "noise_network.", n, c, d, **noise)
    h = a * h
    n = (1.0 - a) * n
    s = h + n
    y = usfgan_residual_blocks(sd, "filter_network.", s, c, d, **filt)
    return _conv_last(sd, y), _conv_last(sd, s), _conv_last(sd, h), _conv_last(sd, n), a


def cascade_hn_usfgan_forward(sd: SD, x: Tensor, c: Tensor, d: Tensor, *, harmonic: dict, noise: dict,
                              filt: dict, upsample_scales: Sequence[int] = (5, 4, 3, 2),
                              pe: Optional[dict] = None):
    """CascadeHnUSFGANGenerator.forward, generator.py:283-334."""
    pe = pe or {}
    c = usfgan_upsample(sd, "upsample_net.", c, upsample_scales)
    a = usfgan_periodicity(sd, "periodicity_estimator.", c, **pe)
    sine, nz = x[:, :1], x[:, 1:2]
    h = conv1x1(sine, fold_weight_norm(sd, "conv_first_sine"), _bias(sd, "conv_first_sine"))
    n = conv1x1(nz, fold_weight_norm(sd, "conv_first_noise"), _bias(sd, "conv_first_noise"))
    h = usfgan_residual_blocks(sd, "harmonic_network.", h, c, d, **harmonic)
    h = a * h
    n = conv1x1(torch.cat([h, n], dim=1), fold_weight_norm(sd, "conv_merge"), _bias(sd, "conv_merge"))
    n = usfgan_residual_blocks(sd, "noise_network.", n, c, d, **noise)
    n = (1.0 - a) * n
    s = h + n
    y = usfgan_residual_blocks(sd, "filter_network.", s, c, d, **filt)
    return _conv_last(sd, y), _conv_last(sd, s), _conv_last(sd, h), _conv_last(sd, n), a


def usfgan_forward(sd: SD, x: Tensor, c: Tensor, d: Tensor, *, source: dict, filt: dict,
                   upsample_scales: Sequence[int] = (5, 4, 3, 2)):
    """USFGANGenerator.forward, generator.py:113-144.  Returns (x, s)."""
    c = usfgan_upsample(sd, "upsample_net.", c, upsample_scales)
    h = conv1x1(x, fold_weight_norm(sd, "conv_first"), _bias(sd, "conv_first"))
    h = usfgan_residual_blocks(sd, "source_network.", h, c, d, **source)
    s = _conv_last(sd, h)
    h = conv1x1(s, fold_weight_norm(sd, "conv_mid"), _bias(sd, "conv_mid"))
    h = usfgan_residual_blocks(sd, "filter_network.", h, c, d, **filt)
    return _conv_last(sd, h), s


# --------------------------------------------------------------------------- #
# uSFGAN front-end (nnsvs/usfgan/utils/features.py, nnsvs/usfgan/__init__.py)
# --------------------------------------------------------------------------- #
def dilated_factor(f0: np.ndarray, fs: int, dense_factor: int) -> np.ndarray:
    """features.py:56-75: fs / (f0 * dense_factor); unvoiced (f0 == 0) -> 1."""
    f0 = np.array(f0, dtype=np.float64, copy=True)
    f0[f0 == 0] = fs / dense_factor
    return fs / f0 / dense_factor


def sine_source(f0: Tensor, sample_rate: int, hop_size: int, sine_amp: float, noise_amp: float,
                noise: Optional[Tensor]) -> Tensor:
    """SignalGenerator.sinusoid with INJECTED noise, features.py:145-164.  f0 (B,1,F)."""
    T = f0.shape[-1] * hop_size
    vuv = F.interpolate((f0 > 0) * torch.ones_like(f0), T)
    rad = (F.interpolate(f0, T) / sample_rate) % 1
    sine = vuv * torch.sin(torch.cumsum(rad, dim=2) * 2 * np.pi) * sine_amp
    if noise_amp > 0:
        amp = vuv * noise_amp + (1.0 - vuv) * noise_amp / 3.0
        sine = sine + noise * amp
    return sine


def usfgan_wrapper_inputs(f0: np.ndarray, aux: Tensor, *, sample_rate: int, hop_size: int, dense_factor: int,
                          aux_context_window: int, sine_amp: float, noise_amp: float,
                          noise_sine: Tensor, noise_in: Tensor):
    """The tensors USFGANWrapper.inference builds before calling the generator
    (nnsvs/usfgan/__init__.py:50-63; signal_types == ["sine", "noise"]), noise injected.
    f0 (F,1) numpy, aux (F,C).  Returns (in_signal (1,2,T), c (1,C,F+2w), d (1,1,T))."""
    df = dilated_factor(np.squeeze(f0.copy()), sample_rate, dense_factor).repeat(hop_size, axis=0)
    c = F.pad(aux.unsqueeze(0).transpose(2, 1), (aux_context_window, aux_context_window), mode="replicate")
    d = torch.FloatTensor(df).view(1, 1, -1)
    f0_t = torch.FloatTensor(f0).unsqueeze(0).transpose(2, 1)
    sine = sine_source(f0_t, sample_rate, hop_size, sine_amp, noise_amp, noise_sine)
    return torch.cat([sine, noise_in], dim=1), c, d


# --------------------------------------------------------------------------- #
# FFConvLSTM encoder  (nnsvs/model.py:779-926) — SURVEY.md §8(f) row 1
# --------------------------------------------------------------------------- #
def lstm_direction(pre: Tensor, w_hh: Tensor, length: int, reverse: bool) -> Tensor:
    """One direction of one ``nn.LSTM`` layer over ONE packed sequence (model.py:917-919).

    pre [T, 4H] = W_ih x_t + b_ih + b_hh; gate rows in torch order i, f, g, o.  The recurrence runs over the first
    ``length`` frames only (``pack_padded_sequence``), from the last of them when ``reverse``; frames beyond stay 0
    (``pad_packed_sequence``).  h_0 = c_0 = 0.
    """
    T, H = pre.shape[0], w_hh.shape[1]
    out = pre.new_zeros((T, H))
    h = pre.new_zeros(H)
    c = pre.new_zeros(H)
    order = range(length - 1, -1, -1) if reverse else range(length)
    for t in order:
        a = pre[t] + w_hh @ h
        i, f, g, o = a[:H], a[H:2 * H], a[2 * H:3 * H], a[3 * H:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return out


def bilstm_stack(sd: SD, prefix: str, x: Tensor, lengths: Sequence[int], num_layers: int) -> Tensor:
    """``nn.LSTM(bidirectional=True, batch_first=True)`` in eval mode on a padded batch x [B, T, C] -> [B, T, 2H]."""
    for layer in range(num_layers):
        outs = []
        for suffix, reverse in (("", False), ("_reverse", True)):
            w_ih = sd[f"{prefix}weight_ih_l{layer}{suffix}"]
            w_hh = sd[f"{prefix}weight_hh_l{layer}{suffix}"]
            b = sd[f"{prefix}bias_ih_l{layer}{suffix}"] + sd[f"{prefix}bias_hh_l{layer}{suffix}"]
            pre = x @ w_ih.t() + b
            outs.append(torch.stack([lstm_direction(pre[i], w_hh, int(lengths[i]), reverse) for i in range(x.shape[0])]))
        x = torch.cat(outs, dim=-1)
    return x


def ffconvlstm_front(sd: SD, x: Tensor, *, in_ph_start_idx: int, in_ph_end_idx: int, embed_dim: Optional[int],
                     spk_embs: Optional[Tensor] = None) -> Tensor:
    """model.py:897-913: phoneme embedding of the one-hot block (argmax, so an all-zero block means phoneme 0) plus a
    Linear over the remaining columns; then the optional speaker embedding is added."""
    if embed_dim is not None:
        s, V = in_ph_start_idx, in_ph_end_idx - in_ph_start_idx
        ph = torch.argmax(x[..., s:s + V], dim=-1)
        rest = torch.cat([x[..., :s], x[..., s + V:]], dim=-1)
        x = sd["emb.weight"][ph] + rest @ sd["fc_in.weight"].t() + sd["fc_in.bias"]
    if spk_embs is not None:
        x = x + spk_embs
    return x


def ffconvlstm_forward(sd: SD, x: Tensor, lengths: Sequence[int], *, in_ph_start_idx: int = 1, in_ph_end_idx: int = 50,
                       embed_dim: Optional[int] = None, num_lstm_layers: int = 2, spk_embs: Optional[Tensor] = None,
                       bn_eps: float = 1e-5, want_parts: bool = False, num_gaussians: int = 4, train: bool = False,
                       bn_momentum: float = 0.1):
    """FFConvLSTM.forward, use_mdn=False (model.py:893-922).  x [B, T, in_dim] -> [B, max(lengths), out_dim].

    ff: 3 x (Linear, ReLU); conv: 3 x (ReflectionPad1d(3), Conv1d k=7, BatchNorm1d, ReLU) over the whole padded batch;
    2-layer BiLSTM over the packed sequences; Linear.  Eval mode: BatchNorm1d with its running statistics.  ``train``
    (module.train() with dropout = 0, what the diffusion recipe's encoders have): BatchNorm1d normalises with the BATCH
    statistics over all B * T positions, padded frames included (biased variance), and moves its running statistics by
    ``bn_momentum`` towards them (unbiased variance) — torch.nn.BatchNorm1d semantics; with want_parts the new buffers are
    returned as parts["bn"] = {"conv.N.running_mean" / "running_var": tensor}.
    """
    x = ffconvlstm_front(sd, x, in_ph_start_idx=in_ph_start_idx, in_ph_end_idx=in_ph_end_idx, embed_dim=embed_dim,
                         spk_embs=spk_embs)
    for i in (0, 2, 4):
        x = torch.relu(x @ sd[f"ff.{i}.weight"].t() + sd[f"ff.{i}.bias"])
    ff = x
    y = x.transpose(1, 2)
    bn_new = {}
    for i in (1, 5, 9):
        y = F.conv1d(F.pad(y, (3, 3), mode="reflect"), sd[f"conv.{i}.weight"], sd[f"conv.{i}.bias"])
        n = i + 1
        if train:
            cnt = y.shape[0] * y.shape[2]
            mean = y.mean(dim=(0, 2))
            var = ((y - mean[None, :, None]) ** 2).mean(dim=(0, 2))
            bn_new[f"conv.{n}.running_mean"] = (1 - bn_momentum) * sd[f"conv.{n}.running_mean"] + bn_momentum * mean
            bn_new[f"conv.{n}.running_var"] = (1 - bn_momentum) * sd[f"conv.{n}.running_var"] + bn_momentum * var * cnt / max(cnt - 1, 1)
        else:
            mean, var = sd[f"conv.{n}.running_mean"], sd[f"conv.{n}.running_var"]
        y = (y - mean[None, :, None]) / torch.sqrt(var[None, :, None] + bn_eps)
        y = torch.relu(y * sd[f"conv.{n}.weight"][None, :, None] + sd[f"conv.{n}.bias"][None, :, None])
    conv = y.transpose(1, 2)
    h = bilstm_stack(sd, "lstm.", conv, lengths, num_lstm_layers)
    h = h[:, :max(int(n) for n in lengths)]
    if "fc.log_pi.weight" in sd:           # use_mdn=True: the head is an MDNLayer (model.py:871-878)
        out = mdn_layer(sd, "fc.", h, num_gaussians)
    else:
        out = h @ sd["fc.weight"].t() + sd["fc.bias"]
    return (out, dict(ff=ff, conv=conv, lstm=h, bn=bn_new)) if want_parts else out


def mdn_layer(sd: SD, prefix: str, h: Tensor, num_gaussians: int) -> Tuple[Tensor, Tensor, Tensor]:
    """MDNLayer.forward with dim_wise=True (nnsvs/mdn.py:45-74): three Linears viewed as [B, T, G, D]; log_softmax of the
    mixture weights over the G components of every output dimension."""
    B, T = h.shape[0], h.shape[1]
    def lin(name):
        return (h @ sd[prefix + name + ".weight"].t() + sd[prefix + name + ".bias"]).view(B, T, num_gaussians, -1)
    return torch.log_softmax(lin("log_pi"), dim=2), lin("log_sigma"), lin("mu")


def mdn_most_probable_sigma_and_mu(log_pi: Tensor, log_sigma: Tensor, mu: Tensor) -> Tuple[Tensor, Tensor]:
    """mdn_get_most_probable_sigma_and_mu for dim-wise mixtures (mdn.py:165-212): per output dimension the component with
    the largest weight; returns (exp(log_sigma), mu) of that component, [B, T, D] each."""
    idx = torch.argmax(log_pi, dim=2, keepdim=True)
    return torch.exp(torch.gather(log_sigma, 2, idx))[:, :, 0], torch.gather(mu, 2, idx)[:, :, 0]


# --------------------------------------------------------------------------- #
# Acoustic post-processing between the diffusion models and the vocoder — SURVEY.md §8(f) row 3
# (nnsvs/postfilters.py:9-46, nnsvs/dsp.py:10-33, call sites nnsvs/gen.py:1394-1418,1500-1513)
# --------------------------------------------------------------------------- #
def variance_scaling(gv: np.ndarray, feats: np.ndarray, offset: int = 2, note_frame_indices=None) -> np.ndarray:
    """postfilters.py:9-46: per-dimension mean / population variance over the note frames of ONE utterance, then
    ``sqrt(gv / utt_gv) * (x - mu) + mu`` on those frames for the dimensions >= offset; no note frames -> unchanged."""
    if note_frame_indices is not None and len(note_frame_indices) == 0:
        return feats
    sel = feats if note_frame_indices is None else feats[note_frame_indices]
    mu, var = sel.mean(0), sel.var(0)
    out = feats.copy()
    rows = slice(None) if note_frame_indices is None else note_frame_indices
    out[rows, offset:] = np.sqrt(gv[offset:] / var[offset:]) * (sel[:, offset:] - mu[offset:]) + mu[offset:]
    return out


def butter_lowpass(N: int, Wn: float) -> Tuple[np.ndarray, np.ndarray]:
    """``scipy.signal.butter(N, Wn, "lowpass")`` (third-party dependency of dsp.py:25; scipy's published algorithm):
    analog Butterworth prototype -> frequency warp -> bilinear transform (fs = 2) -> polynomial coefficients."""
    m = np.arange(-N + 1, N, 2)
    p = -np.exp(1j * np.pi * m / (2 * N))
    warped = 4.0 * np.tan(np.pi * Wn / 2.0)
    p = warped * p
    k = warped ** N
    fs2 = 4.0
    pd = (fs2 + p) / (fs2 - p)
    kd = k * np.real(1.0 / np.prod(fs2 - p))
    b = kd * np.poly(-np.ones(N))
    a = np.real(np.poly(pd))
    return b, a


def lfilter_zi(b: np.ndarray, a: np.ndarray) -> np.ndarray:
    """``scipy.signal.lfilter_zi``: the direct-form-II-transposed state of the step response's steady state,
    (I - A^T) zi = b[1:] - a[1:] b[0] with A the companion matrix of a (a[0] == 1)."""
    n = len(a) - 1
    comp = np.zeros((n, n))
    comp[0] = -a[1:]
    comp[1:, :-1] = np.eye(n - 1)
    return np.linalg.solve(np.eye(n) - comp.T, b[1:] - a[1:] * b[0])


def lfilter_df2t(b: np.ndarray, a: np.ndarray, x: np.ndarray, zi: np.ndarray) -> np.ndarray:
    """``scipy.signal.lfilter(b, a, x, zi=zi)[0]``: y = b0 x + z0; z_i = b_{i+1} x - a_{i+1} y + z_{i+1}."""
    n = len(a) - 1
    z = zi.astype(np.float64).copy()
    y = np.empty(len(x))
    for t, xt in enumerate(x):
        yt = b[0] * xt + z[0]
        for i in range(n - 1):
            z[i] = b[i + 1] * xt - a[i + 1] * yt + z[i + 1]
        z[n - 1] = b[n] * xt - a[n] * yt
        y[t] = yt
    return y


def filtfilt(b: np.ndarray, a: np.ndarray, x: np.ndarray) -> np.ndarray:
    """``scipy.signal.filtfilt(b, a, x)`` with its defaults (dsp.py:31): odd extension by 3 * max(len(a), len(b))
    samples, forward pass started from zi * ext[0], backward pass started from zi * y[-1], extension removed."""
    pad = 3 * max(len(a), len(b))
    x = np.asarray(x, dtype=np.float64)
    ext = np.concatenate([2 * x[0] - x[pad:0:-1], x, 2 * x[-1] - x[-2:-pad - 2:-1]])
    zi = lfilter_zi(b, a)
    y = lfilter_df2t(b, a, ext, zi * ext[0])
    y = lfilter_df2t(b, a, y[::-1], zi * y[-1])[::-1]
    return np.ascontiguousarray(y[pad:-pad])


def lowpass_filter(x: np.ndarray, fs: int, cutoff: float = 5, N: int = 5) -> np.ndarray:
    """dsp.py:10-33: zero-phase Butterworth low-pass of one trajectory; sequences of at most
    max(len(a), len(b)) * (N // 2 + 1) samples come back unchanged."""
    b, a = butter_lowpass(N, cutoff / (fs // 2))
    if len(x) <= max(len(a), len(b)) * (N // 2 + 1):
        return x
    return filtfilt(b, a, x)


def standard_scaler(x: np.ndarray, mean: np.ndarray, scale: np.ndarray, inverse: bool) -> np.ndarray:
    """nnsvs/util.py:288-292 (StandardScaler.transform / inverse_transform).  PARITY UNPINNED for this two-line
    arithmetic: nnsvs.util cannot be imported in the build container (pyworld, hydra, omegaconf), so no fixture was
    generated from the reference class."""
    return x * scale + mean if inverse else (x - mean) / scale


def minmax_scaler(x: np.ndarray, min_: np.ndarray, scale: np.ndarray, inverse: bool) -> np.ndarray:
    """nnsvs/util.py:335-339 (MinMaxScaler.transform / inverse_transform).  PARITY UNPINNED, as above."""
    return (x - min_) / scale if inverse else scale * x + min_
