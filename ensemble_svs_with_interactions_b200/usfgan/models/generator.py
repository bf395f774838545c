"""uSFGAN generators with the constructor / forward / state_dict contract of ``nnsvs.usfgan.models.generator``
(USFGANGenerator generator.py:20-166, CascadeHnUSFGANGenerator :169-356, ParallelHnUSFGANGenerator :359-544).

``forward(x, c, d)`` returns the same tuples as the reference.  ``wave_only=True`` (what USFGANWrapper.inference uses)
skips the extra ``conv_last`` evaluations of s/h/n that the wrapper throws away (usfgan/__init__.py:63).
"""
from logging import getLogger

import torch
import torch.nn as nn

from ... import ops
from ..layers import Conv1d1x1, ResidualBlocks, effective_weight, upsample
from ..layers.residual_block import PeriodicityEstimator
from ..utils import index_initial  # noqa: F401  (API parity)

logger = getLogger(__name__)
f32 = torch.float32


def _conv1x1(m, x, in_relu=False):
    return ops.conv1d_f32(x, effective_weight(m), m.bias, in_relu=in_relu)


def _conv1x1_expand_ntc(m, x_row):
    """1 -> C pointwise conv of one input channel ([B, T] view) straight to NTC bf16."""
    w = effective_weight(m).reshape(-1).to(f32).contiguous()
    return ops.expand1_bf16(x_row, w, m.bias.to(f32).contiguous() if m.bias is not None else None, w.numel())


class _GeneratorBase(nn.Module):
    precision = "auto"   # "auto" | "bf16" | "fp32": bf16 = fused tcgen05 block kernel for the residual stacks

    def _stacks(self):
        return [m for m in self.children() if isinstance(m, ResidualBlocks)]

    def resolved_precision(self):
        ok = all(st.supports_bf16() for st in self._stacks())
        if self.precision == "auto":
            return "bf16" if ok else "fp32"
        if self.precision == "bf16" and not ok:
            raise RuntimeError("uSFGAN precision='bf16' needs residual 64 / gate 128 channels, kernel size 3, aux <= 320")
        if self.precision not in ("bf16", "fp32"):
            raise RuntimeError(f"unknown precision {self.precision!r}")
        return self.precision

    def _run_stack(self, stack, x, c, d, cache, auxb):
        """x NCT fp32 in / out; the stack itself runs NTC bf16 when auxb is given."""
        if auxb is None or len(stack.conv_dilated) == 0:
            return stack(x, c, d, idx_cache=cache)
        xb, _ = ops.nct_to_ntc(x.contiguous())
        yb = stack.forward_ntc_bf16(xb, auxb, d, cache)
        return ops.ntc_bf16_to_nct_f32(yb, x.shape[1])

    def _aux_ntc(self, c):
        if self.resolved_precision() != "bf16":
            return None
        A = c.shape[1]
        auxb, _ = ops.nct_to_ntc(c, Cp=(A + 7) // 8 * 8)
        return auxb

    # "auto": the blocks take the aux projection at FRAME rate (csrc/usfgan_fr.cuh) whenever a 128-sample tile fits the
    # 16-frame window (any hop >= 54 with the recipes' scale lists); "samples": always stream the upsampled features
    aux_projection = "auto"

    def _aux_frames(self, c, T, stacks, cin=None):
        """ops.UsfganAuxFrames for ``stacks`` (block index = position in the concatenation of their blocks), or None when
        the sample-rate path has to be used.  conv1x1_aux(upsample(conv_in(c))) = U . (conv1x1_aux(conv_in(c))): Q is one
        bf16 GEMM at frame rate for all blocks, U the upsampler's impulse responses (cached per frame count)."""
        up = self.upsample_net
        scales = list(up.upsample.upsample_scales)
        hop, reach, rate = 1, 0, 1
        for s_ in scales:
            hop *= s_
        for s_ in scales:
            rate *= s_
            reach += s_ * (hop // rate)
        if self.aux_projection == "samples" or not ops.usfgan_frame_window_ok(hop, reach) or not stacks:
            return None
        if self.aux_projection not in ("auto", "frames"):
            raise RuntimeError(f"unknown aux_projection {self.aux_projection!r}")
        if cin is None:
            cin = up.conv_in_frames(c)                              # [B, A, Tf] fp32
        B, A, Tf = cin.shape
        if Tf * hop != T:
            raise RuntimeError(f"aux features give {Tf} x {hop} samples, the source signal has {T}")
        Ap = (A + 7) // 8 * 8
        cinb, _ = ops.nct_to_ntc(cin, Cp=Ap)
        wkey = tuple((p.data_ptr(), p._version) for st in stacks for b in st.conv_dilated for p in b.conv1x1_aux.parameters())
        if getattr(self, "_auxw_key", None) != wkey:
            self._auxw = torch.cat([st.aux_weights(Ap) for st in stacks], dim=0).to(torch.bfloat16).contiguous()
            self._auxw_key = wkey
        q, fpad = ops.usfgan_aux_frames(cinb, self._auxw, Tf, T, hop, reach)
        ukey = (Tf, str(c.device)) + tuple((p.data_ptr(), p._version) for p in up.upsample.parameters())
        if getattr(self, "_auxu_key", None) != ukey:
            f_idx = torch.arange(Tf, device=c.device)
            impulses = (f_idx[None, :] % 16 == torch.arange(16, device=c.device)[:, None]).to(f32)[None]
            self._auximp = up.upsample(impulses)[0].contiguous()        # [16, T] fp32, kept for _aux_ntc_from_frames
            self._auxu = ops.usfgan_aux_weights(self._auximp, hop, reach)
            self._auxu_key = ukey
        return ops.UsfganAuxFrames(self._auxu, q, fpad, hop, reach)

    def _aux_ntc_from_frames(self, cin, frames, T):
        """Sample-rate aux features [B, T, A8] bf16 (the periodicity estimator's input) as U . conv_in(c) with the impulse
        responses _aux_frames has just cached: one pass at the tensor's write rate (svsk_upsample_frames_bf16)."""
        return ops.upsample_frames_bf16(self._auximp, cin, T, frames.hop, frames.reach)

    def _build_common(self, in_channels, out_channels, residual_channels, skip_channels, aux_channels,
                      aux_context_window, upsample_params):
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.aux_channels = aux_channels
        self.n_ch = residual_channels
        self.upsample_net = getattr(upsample, "ConvInUpsampleNetwork")(
            **upsample_params, aux_channels=aux_channels, aux_context_window=aux_context_window)

    def _make_conv_last(self, skip_channels, out_channels):
        return nn.Sequential(nn.ReLU(), Conv1d1x1(skip_channels, skip_channels), nn.ReLU(),
                             Conv1d1x1(skip_channels, out_channels))

    def _conv_last(self, x):
        x = _conv1x1(self.conv_last[1], x, in_relu=True)
        return _conv1x1(self.conv_last[3], x, in_relu=True)

    def _conv_last_ntc_bf16(self, yb_relu):
        """conv_last on an NTC bf16 tensor that already went through the leading ReLU: 1x1 (tensor cores, ReLU fused)
        then the C -> 1 projection.  Returns (B, 1, T) fp32."""
        l1, l3 = self.conv_last[1], self.conv_last[3]
        key = tuple((p.data_ptr(), p._version) for p in self.conv_last.parameters())
        if getattr(self, "_last_key", None) != key:
            with torch.no_grad():
                self._last_plan = (ops.conv1d_pack_bf16(effective_weight(l1).to(f32).contiguous()),
                                   l1.bias.detach().to(f32).contiguous(),
                                   effective_weight(l3).to(f32).reshape(-1).contiguous(), float(l3.bias.detach()[0]))
            self._last_key = key
        wp, b1, w2, b2 = self._last_plan
        hb = ops.conv1d_bf16(yb_relu, wp, b1, l1.out_channels, 1, act=ops.ACT_RELU)
        return ops.dot_rows_bf16(hb, w2, b2).unsqueeze(1)

    @staticmethod
    def _check(x):
        if not x.is_cuda:
            raise RuntimeError("uSFGAN generators run on CUDA (sm_100a) only: libsvsk has no CPU path")

    def remove_weight_norm(self):
        """Remove weight normalization from all layers (generator.py:146-155)."""

        def _remove(m):
            try:
                nn.utils.remove_weight_norm(m)
            except ValueError:
                return

        self.apply(_remove)

    def apply_weight_norm(self):
        """Apply old-style weight normalization to every Conv1d / Conv2d holder (generator.py:157-166)."""

        def _apply(m):
            if isinstance(m, (nn.Conv1d, nn.Conv2d)):
                nn.utils.weight_norm(m)

        self.apply(_apply)


def _with_widths(params, residual_channels, gate_channels, skip_channels, aux_channels):
    p = dict(params)
    p.update(residual_channels=residual_channels, gate_channels=gate_channels, skip_channels=skip_channels,
             aux_channels=aux_channels)
    return p


class USFGANGenerator(_GeneratorBase):
    def __init__(self,
                 source_network_params={"blockA": 30, "cycleA": 3, "blockF": 0, "cycleF": 0, "cascade_mode": 0},
                 filter_network_params={"blockA": 0, "cycleA": 0, "blockF": 30, "cycleF": 3, "cascade_mode": 0},
                 in_channels=1, out_channels=1, residual_channels=64, gate_channels=128, skip_channels=64,
                 aux_channels=80, aux_context_window=2, use_weight_norm=True,
                 upsample_params={"upsample_scales": [5, 4, 3, 2]}):
        super().__init__()
        self.conv_first = Conv1d1x1(in_channels, residual_channels)
        self._build_common(in_channels, out_channels, residual_channels, skip_channels, aux_channels,
                           aux_context_window, upsample_params)
        widths = (residual_channels, gate_channels, skip_channels, aux_channels)
        self.source_network = ResidualBlocks(**_with_widths(source_network_params, *widths))
        self.filter_network = ResidualBlocks(**_with_widths(filter_network_params, *widths))
        self.conv_mid = Conv1d1x1(out_channels, skip_channels)
        self.conv_last = self._make_conv_last(skip_channels, out_channels)
        if use_weight_norm:
            self.apply_weight_norm()

    @torch.no_grad()
    def forward(self, x, c, d):
        """x (B,1,T), c (B,C,T'), d (B,1,T) -> (x, s)   (generator.py:113-144)."""
        self._check(x)
        cache = {}
        l1, l3 = self.conv_last[1], self.conv_last[3]
        if (self.resolved_precision() == "bf16" and self.in_channels == 1 and l3.out_channels == 1 and l1.in_channels % 8 == 0
                and l1.out_channels % 16 == 0):
            # everything stays NTC bf16 (as in ParallelHnUSFGANGenerator's wave-only path): aux features upsampled straight
            # into that layout, the 1 -> C convs write it directly, conv_last runs on the tensor cores
            frames = self._aux_frames(c, x.size(-1), [self.source_network, self.filter_network])
            auxb = None
            if frames is None:
                if self.upsample_net.supports_fused():
                    auxb = self.upsample_net.forward_ntc_bf16(c)
                else:
                    auxb = self._aux_ntc(self.upsample_net(c))
                assert auxb.size(1) == x.size(-1)
            xb = _conv1x1_expand_ntc(self.conv_first, x.to(f32).contiguous()[:, 0])
            yb = self.source_network.forward_ntc_bf16(xb, auxb, d, cache, relu_last=True, frames=frames)
            s = self._conv_last_ntc_bf16(yb)
            xb = _conv1x1_expand_ntc(self.conv_mid, s[:, 0])
            yb = self.filter_network.forward_ntc_bf16(xb, auxb, d, cache, relu_last=True, frames=frames,
                                                      frames_block0=len(self.source_network.conv_dilated))
            return self._conv_last_ntc_bf16(yb), s
        c = self.upsample_net(c)
        assert c.size(-1) == x.size(-1)
        auxb = self._aux_ntc(c)
        x = _conv1x1(self.conv_first, x.to(f32).contiguous())
        x = self._run_stack(self.source_network, x, c, d, cache, auxb)
        s = self._conv_last(x)
        x = _conv1x1(self.conv_mid, s)
        x = self._run_stack(self.filter_network, x, c, d, cache, auxb)
        return self._conv_last(x), s


class _HnBase(_GeneratorBase):
    supports_wave_only = True
    has_merge = False

    def _build_hn(self, harmonic_network_params, noise_network_params, filter_network_params,
                  periodicity_estimator_params, in_channels, out_channels, residual_channels, gate_channels,
                  skip_channels, aux_channels, aux_context_window, upsample_params):
        # registration order == the reference's, so state_dict key order matches too
        self.conv_first_sine = Conv1d1x1(in_channels, residual_channels)
        self.conv_first_noise = Conv1d1x1(in_channels, residual_channels)
        if self.has_merge:
            self.conv_merge = Conv1d1x1(residual_channels * 2, residual_channels)
        self._build_common(in_channels, out_channels, residual_channels, skip_channels, aux_channels,
                           aux_context_window, upsample_params)

    def _build_networks(self, harmonic_network_params, noise_network_params, filter_network_params,
                        periodicity_estimator_params, residual_channels, gate_channels, skip_channels, aux_channels,
                        out_channels):
        widths = (residual_channels, gate_channels, skip_channels, aux_channels)
        self.harmonic_network = ResidualBlocks(**_with_widths(harmonic_network_params, *widths))
        self.noise_network = ResidualBlocks(**_with_widths(noise_network_params, *widths))
        self.filter_network = ResidualBlocks(**_with_widths(filter_network_params, *widths))
        # NB like the reference (generator.py:453-455) the estimator's width is NOT tied to residual_channels:
        # it is 64 unless periodicity_estimator_params carries "residual_channels".
        self.periodicity_estimator = PeriodicityEstimator(**periodicity_estimator_params, in_channels=aux_channels)
        self.conv_last = self._make_conv_last(skip_channels, out_channels)

    def _front(self, x, c):
        c = self.upsample_net(c)
        assert c.size(-1) == x.size(-1)
        a = self.periodicity_estimator(c)
        x = x.to(f32)
        sine, noise = x[:, 0:1].contiguous(), x[:, 1:2].contiguous()
        return c, a, _conv1x1(self.conv_first_sine, sine), _conv1x1(self.conv_first_noise, noise)

    def _ntc_fast_path_ok(self):
        l1, l3 = self.conv_last[1], self.conv_last[3]
        pe_ok = self.periodicity_estimator.supports_bf16() and \
            self.periodicity_estimator.layers[2 * (self.periodicity_estimator.conv_layers - 1)].out_channels == self.n_ch
        return (self.resolved_precision() == "bf16" and pe_ok and l3.out_channels == 1 and l1.in_channels % 8 == 0
                and l1.out_channels % 16 == 0 and self.in_channels == 1)

    def _outputs(self, y, s, h, n, a, wave_only):
        y = self._conv_last(y)
        if wave_only:
            return y, None, None, None, a
        return y, self._conv_last(s), self._conv_last(h), self._conv_last(n), a


_DEFAULT_HARMONIC = {"blockA": 20, "cycleA": 4, "blockF": 0, "cycleF": 0, "cascade_mode": 0}
_DEFAULT_NOISE = {"blockA": 0, "cycleA": 0, "blockF": 5, "cycleF": 5, "cascade_mode": 0}
_DEFAULT_FILTER = {"blockA": 0, "cycleA": 0, "blockF": 30, "cycleF": 3, "cascade_mode": 0}
_DEFAULT_PE = {"conv_blocks": 3, "kernel_size": 5, "dilation": 1, "padding_mode": "replicate"}


class CascadeHnUSFGANGenerator(_HnBase):
    has_merge = True

    def __init__(self, harmonic_network_params=_DEFAULT_HARMONIC, noise_network_params=_DEFAULT_NOISE,
                 filter_network_params=_DEFAULT_FILTER, periodicity_estimator_params=_DEFAULT_PE, in_channels=1,
                 out_channels=1, residual_channels=64, gate_channels=128, skip_channels=64, aux_channels=80,
                 aux_context_window=2, use_weight_norm=True, upsample_params={"upsample_scales": [5, 4, 3, 2]}):
        super().__init__()
        self._build_hn(harmonic_network_params, noise_network_params, filter_network_params,
                       periodicity_estimator_params, in_channels, out_channels, residual_channels, gate_channels,
                       skip_channels, aux_channels, aux_context_window, upsample_params)
        self._build_networks(harmonic_network_params, noise_network_params, filter_network_params,
                             periodicity_estimator_params, residual_channels, gate_channels, skip_channels,
                             aux_channels, out_channels)
        if use_weight_norm:
            self.apply_weight_norm()

    @torch.no_grad()
    def forward(self, x, c, d, wave_only=False):
        """(x, s, h, n, a)   (generator.py:283-334): harmonic -> a*h -> merge with noise input -> noise net."""
        self._check(x)
        cache = {}
        if wave_only and self._ntc_fast_path_ok() and self.conv_merge.in_channels % 8 == 0 and self.conv_merge.out_channels % 16 == 0:
            # NTC bf16 throughout, like ParallelHnUSFGANGenerator's wave-only path.  s = a h + (1 - a) n is formed from the
            # two stacks' raw outputs in one pass; a h alone is only needed as the merge conv's input.
            cin = self.upsample_net.conv_in_frames(c)   # once: the periodicity estimator's input and the blocks' Q share it
            nets = [self.harmonic_network, self.noise_network, self.filter_network]
            frames = self._aux_frames(c, x.size(-1), nets, cin=cin)
            if frames is not None and cin.shape[1] <= 112:
                auxb = self._aux_ntc_from_frames(cin, frames, x.size(-1))
            elif self.upsample_net.supports_fused():
                auxb = self.upsample_net.forward_ntc_bf16(c, cin=cin)
            else:
                auxb = self._aux_ntc(self.upsample_net.upsample(cin))
            assert auxb.size(1) == x.size(-1)
            ab = self.periodicity_estimator.forward_ntc_bf16(auxb)
            xf = x.to(f32).contiguous()
            hb = _conv1x1_expand_ntc(self.conv_first_sine, xf[:, 0])
            nb = _conv1x1_expand_ntc(self.conv_first_noise, xf[:, 1])
            first = [0, len(nets[0].conv_dilated), len(nets[0].conv_dilated) + len(nets[1].conv_dilated)]
            hb = self.harmonic_network.forward_ntc_bf16(hb, auxb, d, cache, frames=frames, frames_block0=first[0])
            zeros = torch.zeros_like(hb)
            hm = ops.periodic_mix_bf16(ab, hb, zeros)                       # a * h
            key = tuple((p.data_ptr(), p._version) for p in self.conv_merge.parameters())
            if getattr(self, "_merge_key", None) != key:
                self._merge_plan = (ops.conv1d_pack_bf16(effective_weight(self.conv_merge).to(f32).contiguous()),
                                    self.conv_merge.bias.detach().to(f32).contiguous())
                self._merge_key = key
            nb = ops.conv1d_bf16(torch.cat([hm, nb], dim=2), self._merge_plan[0], self._merge_plan[1], self.conv_merge.out_channels, 1)
            nb = self.noise_network.forward_ntc_bf16(nb, auxb, d, cache, frames=frames, frames_block0=first[1])
            sb = ops.periodic_mix_bf16(ab, hb, nb)                          # a * h + (1 - a) * n
            yb = self.filter_network.forward_ntc_bf16(sb, auxb, d, cache, relu_last=True, frames=frames, frames_block0=first[2])
            return self._conv_last_ntc_bf16(yb), None, None, None, ab
        c, a, h, n = self._front(x, c)
        auxb = self._aux_ntc(c)
        h = self._run_stack(self.harmonic_network, h, c, d, cache, auxb)
        zeros = torch.zeros_like(h)
        _, h, _ = ops.periodic_mix_f32(a, h, zeros, want_parts=True)     # h <- a * h
        n = _conv1x1(self.conv_merge, torch.cat([h, n], dim=1))
        n = self._run_stack(self.noise_network, n, c, d, cache, auxb)
        _, _, n = ops.periodic_mix_f32(a, zeros, n, want_parts=True)     # n <- (1 - a) * n
        s = ops.lincomb_f32([h, n], [1.0, 1.0])
        y = self._run_stack(self.filter_network, s, c, d, cache, auxb)
        return self._outputs(y, s, h, n, a, wave_only)


class ParallelHnUSFGANGenerator(_HnBase):
    def __init__(self, harmonic_network_params=_DEFAULT_HARMONIC, noise_network_params=_DEFAULT_NOISE,
                 filter_network_params=_DEFAULT_FILTER, periodicity_estimator_params=_DEFAULT_PE, in_channels=1,
                 out_channels=1, residual_channels=64, gate_channels=128, skip_channels=64, aux_channels=80,
                 aux_context_window=2, use_weight_norm=True, upsample_params={"upsample_scales": [5, 4, 3, 2]}):
        super().__init__()
        self._build_hn(harmonic_network_params, noise_network_params, filter_network_params,
                       periodicity_estimator_params, in_channels, out_channels, residual_channels, gate_channels,
                       skip_channels, aux_channels, aux_context_window, upsample_params)
        self._build_networks(harmonic_network_params, noise_network_params, filter_network_params,
                             periodicity_estimator_params, residual_channels, gate_channels, skip_channels,
                             aux_channels, out_channels)
        if use_weight_norm:
            self.apply_weight_norm()

    @torch.no_grad()
    def forward(self, x, c, d, wave_only=False):
        """(x, s, h, n, a)   (generator.py:472-522): harmonic || noise -> a*h + (1-a)*n -> filter."""
        self._check(x)
        cache = {}
        if wave_only and self._ntc_fast_path_ok():
            # everything stays NTC bf16: the aux features are upsampled straight into that layout (one pass over all
            # stages), the two 1 -> C input convs write it directly
            cin = self.upsample_net.conv_in_frames(c)   # once: the periodicity estimator's input and the blocks' Q share it
            nets = [self.harmonic_network, self.noise_network, self.filter_network]
            frames = self._aux_frames(c, x.size(-1), nets, cin=cin)
            if frames is not None and cin.shape[1] <= 112:
                auxb = self._aux_ntc_from_frames(cin, frames, x.size(-1))
            elif self.upsample_net.supports_fused():
                auxb = self.upsample_net.forward_ntc_bf16(c, cin=cin)
            else:
                auxb = self._aux_ntc(self.upsample_net.upsample(cin))
            assert auxb.size(1) == x.size(-1)
            ab = self.periodicity_estimator.forward_ntc_bf16(auxb)
            xf = x.to(f32).contiguous()
            hb = _conv1x1_expand_ntc(self.conv_first_sine, xf[:, 0])
            nb = _conv1x1_expand_ntc(self.conv_first_noise, xf[:, 1])
            first = [0, len(nets[0].conv_dilated), len(nets[0].conv_dilated) + len(nets[1].conv_dilated)]
            hb = self.harmonic_network.forward_ntc_bf16(hb, auxb, d, cache, frames=frames, frames_block0=first[0])
            nb = self.noise_network.forward_ntc_bf16(nb, auxb, d, cache, frames=frames, frames_block0=first[1])
            sb = ops.periodic_mix_bf16(ab, hb, nb)
            yb = self.filter_network.forward_ntc_bf16(sb, auxb, d, cache, relu_last=True, frames=frames, frames_block0=first[2])
            return self._conv_last_ntc_bf16(yb), None, None, None, ab
        c, a, h, n = self._front(x, c)
        auxb = self._aux_ntc(c)
        h = self._run_stack(self.harmonic_network, h, c, d, cache, auxb)
        n = self._run_stack(self.noise_network, n, c, d, cache, auxb)
        s, h, n = ops.periodic_mix_f32(a, h, n, want_parts=True)
        y = self._run_stack(self.filter_network, s, c, d, cache, auxb)
        return self._outputs(y, s, h, n, a, wave_only)
