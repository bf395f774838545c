"""Gaussian diffusion wrapper — drop-in for ``nnsvs.diffsinger.diffusion`` (nnsvs/diffsinger/diffusion.py:54-440).

Same constructor kwargs, the same 12 registered schedule buffers (part of ``state_dict``), the same
``forward`` / ``inference`` / ``p_sample`` / ``q_sample`` / ``p_sample_plms`` signatures.  The sampling loop keeps
the DDPM state in the denoiser's native layout for the whole K-step loop (no per-step layout conversion), uses a
step-bias table pre-computed for t = 0..K-1 (the step embedding depends only on t) and can replay the
entire loop as one CUDA graph.
"""
from __future__ import annotations

from collections import deque

import numpy as np
import math
import os

import torch

from .. import _lib as _L
from .. import ops
from ..base import BaseModel, PredictionType

f32 = torch.float32

SCHEDULE_BUFFERS = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
    "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
    "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2",
)


def linear_beta_schedule(timesteps, min_beta=1e-4, max_beta=0.06):
    """diffusion.py:27-32."""
    return np.linspace(min_beta, max_beta, timesteps)


def cosine_beta_schedule(timesteps, s=0.008):
    """diffusion.py:35-45 (https://openreview.net/forum?id=-NEXDKk8gZ)."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    return np.clip(1 - (ac[1:] / ac[:-1]), a_min=0, a_max=0.999)


beta_schedule = {"cosine": cosine_beta_schedule, "linear": linear_beta_schedule}


class GaussianDiffusion(BaseModel):
    def __init__(self, in_dim, out_dim, denoise_fn, encoder=None, K_step=100, betas=None, schedule_type="linear",
                 scheduler_params=None, norm_scale=10, pndm_speedup=None):
        super().__init__()
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.denoise_fn = denoise_fn
        self.K_step = K_step
        self.pndm_speedup = pndm_speedup
        self.encoder = encoder
        self.norm_scale = norm_scale
        if scheduler_params is None:
            scheduler_params = {"max_beta": 0.06} if schedule_type == "linear" else {"s": 0.008}
        if encoder is not None:
            assert encoder.in_dim == in_dim, "encoder input dim must match in_dim"
        assert out_dim == denoise_fn.in_dim, "denoise_fn input dim must match out_dim"
        if pndm_speedup:
            # same contract as the reference constructor (diffusion.py:86-87); the PLMS sampler itself is
            # implemented and reachable by setting the attribute afterwards, as in the reference.
            raise NotImplementedError("pndm_speedup is not implemented yet")

        if betas is not None:
            betas = betas.detach().cpu().numpy() if isinstance(betas, torch.Tensor) else betas
        else:
            betas = beta_schedule[schedule_type](K_step, **scheduler_params)
        betas = np.asarray(betas, dtype=np.float64)
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        pv = betas * (1.0 - ac_prev) / (1.0 - ac)
        self.noise_list = deque(maxlen=4)
        tables = {
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
        for name in SCHEDULE_BUFFERS:  # registration order == the reference's (state_dict key order)
            self.register_buffer(name, torch.tensor(tables[name], dtype=torch.float32))
        self.use_cuda_graph = True
        self._graphs = {}

    # ------------------------------------------------------------------ small API mirrors
    def _norm(self, x, a_max=10):
        return x / a_max

    def _denorm(self, x, a_max=10):
        return x * a_max

    def prediction_type(self):
        return PredictionType.DIFFUSION

    def _tables(self):
        return {k: getattr(self, k) for k in SCHEDULE_BUFFERS}

    def _require_cuda(self, t):
        if not t.is_cuda:
            raise RuntimeError("GaussianDiffusion runs on CUDA (sm_100a) only: libsvsk has no CPU path")

    def q_sample(self, x_start, t, noise=None):
        """diffusion.py:261-267."""
        self._require_cuda(x_start)
        if noise is None:
            noise = torch.randn_like(x_start)
        return ops.q_sample_f32(x_start.to(f32).contiguous(), noise.to(f32).contiguous(), t.to(torch.int64).contiguous(),
                                self._tables())

    @torch.no_grad()
    def p_sample(self, x, t, cond, noise_fn=torch.randn, clip_denoised=True, repeat_noise=False):
        """diffusion.py:193-204 — one ancestral step in the reference's (B,1,M,T) layout."""
        self._require_cuda(x)
        eps = self.denoise_fn(x, t, cond=cond)
        if repeat_noise:
            noise = noise_fn(1, *x.shape[1:], device=x.device).repeat(x.shape[0], *([1] * (x.dim() - 1)))
        else:
            noise = noise_fn(*x.shape, device=x.device)
        return ops.ddpm_update_f32(x.to(f32).contiguous(), eps.contiguous(), noise.to(f32).contiguous(),
                                   t.to(torch.int64).contiguous(), self._tables(), clip_denoised)

    @torch.no_grad()
    def p_sample_plms(self, x, t, interval, cond):
        """diffusion.py:206-259 (PLMS, https://arxiv.org/abs/2202.09778)."""
        self._require_cuda(x)
        x = x.to(f32).contiguous()
        t = t.to(torch.int64).contiguous()
        nl = self.noise_list
        e = self.denoise_fn(x, t, cond=cond).contiguous()
        if len(nl) == 0:
            x_pred = ops.plms_transfer_f32(x, e, t, interval, self.alphas_cumprod)
            e_prev = self.denoise_fn(x_pred, torch.clamp(t - interval, min=0), cond=cond).contiguous()
            e_prime = ops.lincomb_f32([e, e_prev], [0.5, 0.5])
        elif len(nl) == 1:
            e_prime = ops.lincomb_f32([e, nl[-1]], [3 / 2, -1 / 2])
        elif len(nl) == 2:
            e_prime = ops.lincomb_f32([e, nl[-1], nl[-2]], [23 / 12, -16 / 12, 5 / 12])
        else:
            e_prime = ops.lincomb_f32([e, nl[-1], nl[-2], nl[-3]], [55 / 24, -59 / 24, 37 / 24, -9 / 24])
        x_prev = ops.plms_transfer_f32(x, e_prime, t, interval, self.alphas_cumprod)
        nl.append(e)
        return x_prev

    # ------------------------------------------------------------------ training forward (diffusion.py:269-300)
    def forward(self, cond, lengths=None, y=None, spk_embs=None, *, t=None, noise=None):
        """Returns (noise, eps_hat), both (B, T, out_dim).  ``t`` / ``noise`` may be injected (parity tests)."""
        self._require_cuda(cond)
        B = cond.shape[0]
        device = cond.device
        if self.encoder is not None:
            cond = self.encoder(cond, lengths, spk_embs=spk_embs)
        cond = cond.transpose(1, 2)
        if t is None:
            t = torch.randint(0, self.K_step, (B,), device=device).long()
        x = self._norm(y, self.norm_scale)
        x = x.transpose(1, 2)[:, None, :, :]
        if noise is None:
            noise = torch.randn_like(x)
        x_noisy = self.q_sample(x_start=x.contiguous(), t=t, noise=noise.contiguous())
        x_recon = self.denoise_fn(x_noisy, t, cond)
        return noise.squeeze(1).transpose(1, 2), x_recon.squeeze(1).transpose(1, 2)

    # ------------------------------------------------------------------ sampling (diffusion.py:302-336)
    def _step_table(self):
        """[L, K, 3*2C] tap biases for t = 0..K-1 (cached with the denoiser's packed weights)."""
        den = self.denoise_fn
        plan = den.bf16_plan()
        if plan.step_table is None or plan.step_table.shape[1] != self.K_step:
            t_all = torch.arange(self.K_step, device=self.betas.device, dtype=torch.int64)
            plan.step_table = den.step_bias_bf16(t_all)
        return plan.step_table

    def _sample_loop_bf16(self, condb, x32s, z_ntc, trace=None):
        """x32s [B,T,Mp] fp32 (updated in place), z_ntc [K,B,T,Mp] fp32.  ``trace``: {t: None} -> filled with the
        denoiser output of step t, [B,T,Mp] fp32 (parity harness; eager launches only)."""
        den = self.denoise_fn
        plan = den.bf16_plan()
        table = self._step_table()
        B = x32s.shape[0]
        tabs = self._tables()
        if os.environ.get("SVSK_DIFFNET_STEP", "1") != "0" and plan.Mp <= 128:
            # per step: the residual stack, then ONE kernel for tail projections + p_sample update + the next step's
            # input projection (svsk_diffnet_step_bf16)
            sched = [tabs[k] for k in ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
                                       "posterior_mean_coef2", "posterior_log_variance_clipped")]
            xb0 = den.project_in_bf16(x32s, plan)
            # cond is the same for all K calls: its projection through every layer's conditioner_projection once per run
            pcond = den.cond_projection_bf16(condb, plan) if self.K_step > 1 else None
            for i in reversed(range(self.K_step)):
                skip32 = den.residual_stack_bf16(xb0, condb, table[:, i:i + 1], plan, pcond=pcond)   # row i of every layer's table
                eps_out = torch.empty_like(x32s) if trace is not None and i in trace else None
                ops.diffnet_step_bf16(skip32, x32s, z_ntc[i], self._t_const[i], sched, plan.w_skip, plan.b_skip, plan.w_out,
                                      plan.b_out, skip_scale=1.0 / math.sqrt(plan.L), w_in=plan.w_in, b_in=plan.b_in,
                                      xb_out=xb0 if i > 0 else None, eps_out=eps_out)
                if eps_out is not None:
                    trace[i] = eps_out
            return x32s
        for i in reversed(range(self.K_step)):
            # row i of every layer's table, shared by the whole batch
            eps = den.denoise_ntc_bf16(x32s, condb, table[:, i:i + 1], plan=plan)
            if trace is not None and i in trace:
                trace[i] = eps.clone()
            ops.ddpm_update_f32(x32s, eps, z_ntc[i], self._t_const[i], tabs, True, out=x32s)
        return x32s

    def _sample_loop_fp32(self, cond_nct, x, z, trace=None):
        den = self.denoise_fn
        tabs = self._tables()
        for i in reversed(range(self.K_step)):
            t = self._t_const[i]
            eps = den.denoise_nct_fp32(x[:, 0], cond_nct, t)[:, None]
            if trace is not None and i in trace:
                trace[i] = eps.clone()
            ops.ddpm_update_f32(x, eps.contiguous(), z[i], t, tabs, True, out=x)
        return x

    def _prepare_t_const(self, B, device):
        # One [B] int64 tensor per step (views of a [K,B] table): no per-step host->device traffic.  The tables are kept
        # per (K, B, device) for the module's lifetime: captured CUDA graphs hold their device addresses, so a table
        # must not be freed when a call with another batch size comes in between two replays.
        key = (self.K_step, B, str(device))
        cache = self.__dict__.setdefault("_t_const_cache", {})
        if key not in cache:
            tt = torch.arange(self.K_step, device=device, dtype=torch.int64)[:, None].expand(self.K_step, B).contiguous()
            cache[key] = [tt[i] for i in range(self.K_step)]
        self._t_const = cache[key]

    @torch.no_grad()
    def sample(self, cond_t, x_T=None, z=None, trace=None):
        """Core of ``inference`` after the encoder.  cond_t (B,H,T) fp32; optional injected x_T (B,1,M,T) and
        z (K,B,1,M,T) (z[i] is consumed at step t=i).  Returns (B,T,M) * norm_scale.
        ``trace`` (parity harness): a dict whose keys are step indices t; on return trace[t] is the denoiser output of
        that step as (B,1,M,T) fp32.  Tracing runs the eager launches (bit-identical to the graph replay)."""
        den = self.denoise_fn
        B, H, T = cond_t.shape
        M = self.out_dim
        device = cond_t.device
        cond_t = cond_t.to(f32).contiguous()
        self._prepare_t_const(B, device)
        if den.resolved_precision() == "fp32":
            x = torch.randn((B, 1, M, T), device=device) if x_T is None else x_T.to(f32).clone()
            zz = torch.randn((self.K_step, B, 1, M, T), device=device) if z is None else z.to(f32).contiguous()
            x = self._sample_loop_fp32(cond_t, x.contiguous(), zz, trace)
            return (x[:, 0].transpose(1, 2) * self.norm_scale).contiguous()

        plan = den.bf16_plan()
        Mp = plan.Mp
        condb, _ = ops.nct_to_ntc(cond_t)
        x32s = z_ntc = None
        if x_T is not None:
            _, x32s = ops.nct_to_ntc(x_T[:, 0].to(f32).contiguous(), Cp=Mp, want_bf16=False, want_f32=True)
        if z is not None:
            zf = z.to(f32).reshape(self.K_step * B, M, T).contiguous()
            _, z_ntc = ops.nct_to_ntc(zf, Cp=Mp, want_bf16=False, want_f32=True)
            z_ntc = z_ntc.view(self.K_step, B, T, Mp)
        self._step_table()  # make sure the table exists before a graph capture
        if self.use_cuda_graph and trace is None:
            x32s = self._graph_replay(condb, x32s, z_ntc, (B, T, Mp))
        else:
            # same draw order as the graph path: x_T first, then the K noise tensors
            if x32s is None:
                x32s = torch.randn((B, T, Mp), device=device)
            if z_ntc is None:
                z_ntc = torch.randn((self.K_step, B, T, Mp), device=device)
            x32s = self._sample_loop_bf16(condb, x32s, z_ntc, trace)
            if trace is not None:
                for k in list(trace):
                    if trace[k] is not None:
                        trace[k] = trace[k][:, :, :M].transpose(1, 2)[:, None].contiguous()
        out = x32s[:, :, :M] * self.norm_scale
        return out.contiguous()

    def _graph_replay(self, condb, x32s, z_ntc, xshape):
        """Capture the whole K-step loop once per (B,T) and replay it (static buffers, ~26 K launches -> 1).
        ``x32s`` / ``z_ntc`` None: the Gaussian draws go straight into the graph's static buffers (no 384 MB staging
        tensor and copy per pass at BASELINE config 2)."""
        zshape = (self.K_step,) + tuple(xshape)
        key = (tuple(condb.shape), tuple(xshape), self.K_step, self.denoise_fn._param_key())
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= 4:
                self._graphs.clear()
            s_cond = torch.empty_like(condb)
            s_x = torch.zeros(xshape, device=condb.device, dtype=f32)
            s_z = torch.zeros(zshape, device=condb.device, dtype=f32)
            s_cond.copy_(condb)
            # warm-up outside capture (function attributes, tensor-map cache, allocator pools)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._sample_loop_bf16(s_cond, s_x.clone(), s_z)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            n0 = _L.launch_count
            with torch.cuda.graph(g):
                self._sample_loop_bf16(s_cond, s_x, s_z)
            ent = (g, s_cond, s_x, s_z, _L.launch_count - n0)
            _L.launch_count = n0  # capture enqueues nothing; replays are counted below
            self._graphs[key] = ent
        g, s_cond, s_x, s_z, n_kernels = ent
        s_cond.copy_(condb)
        if x32s is None:
            s_x.normal_()
        else:
            s_x.copy_(x32s)
        if z_ntc is None:
            s_z.normal_()
        else:
            s_z.copy_(z_ntc)
        g.replay()
        _L.launch_count += n_kernels
        return s_x

    @torch.no_grad()
    def inference(self, cond, lengths=None, spk_embs=None, *, x_T=None, z=None, cond_is_encoded=False, trace=None):
        """diffusion.py:296-336.  ``cond_is_encoded``: ``cond`` already is the encoder's output (pipeline.py runs the
        encoders of several streams side by side), so the encoder is skipped."""
        self._require_cuda(cond)
        B = cond.shape[0]
        device = cond.device
        if self.encoder is not None and not cond_is_encoded:
            cond = self.encoder(cond, lengths, spk_embs=spk_embs)
        cond = cond.transpose(1, 2)  # (B, H, T)
        if self.pndm_speedup:
            t = self.K_step
            x = torch.randn((B, 1, self.out_dim, cond.shape[2]), device=device) if x_T is None else x_T
            self.noise_list = deque(maxlen=4)
            interval = int(self.pndm_speedup)
            cond_c = cond.contiguous()
            for i in reversed(range(0, t, interval)):
                x = self.p_sample_plms(x, torch.full((B,), i, device=device, dtype=torch.long), interval, cond_c)
            return self._denorm(x[:, 0].transpose(1, 2), self.norm_scale)
        return self.sample(cond, x_T=x_T, z=z, trace=trace)


class MultiSpeakerGaussianDiffusion(GaussianDiffusion):
    """diffusion.py:339-440: speaker embedding expanded over time and handed to the encoder."""

    def __init__(self, in_dim, out_dim, denoise_fn, speaker_embedding, encoder=None, K_step=100, betas=None,
                 schedule_type="linear", scheduler_params=None, norm_scale=10, pndm_speedup=None):
        super().__init__(in_dim, out_dim, denoise_fn, encoder, K_step, betas, schedule_type, scheduler_params,
                         norm_scale, pndm_speedup)
        self.speaker_embedding = speaker_embedding

    def _spk(self, cond, spks):
        e = self.speaker_embedding(spks)
        return e.expand(e.shape[0], cond.shape[1], e.shape[-1])

    def forward(self, cond, spks, lengths=None, y=None):
        return super().forward(cond, lengths, y, spk_embs=self._spk(cond, spks))

    def inference(self, cond, spks, lengths=None):
        return super().inference(cond, lengths, spk_embs=self._spk(cond, spks))
