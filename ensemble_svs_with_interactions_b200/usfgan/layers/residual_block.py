"""QPPWG / uSFGAN residual blocks with the module tree of ``nnsvs.usfgan.layers.residual_block``
(residual_block.py:27-399): ``FixedBlock``, ``AdaptiveBlock``, ``ResidualBlocks``, ``PeriodicityEstimator``.

The ``nn.Conv1d`` sub-modules hold the parameters (incl. old-style weight-norm ``weight_g`` / ``weight_v``), so
vocoder checkpoints load with strict=True before or after ``remove_weight_norm()``.  Forward arithmetic is libsvsk:

* fp32: svsk_conv1d_f32 (REFLECT / INDEXED taps) + svsk_gated_act_f32, residual add and sqrt(1/2) scale fused into the
  output projection's epilogue;
* ``conv1x1_skip`` is never evaluated: ``ResidualBlocks.forward`` of the reference discards the skip sum
  (residual_block.py:333-336) — the parameters stay in the state_dict, the work is skipped.
"""
import math

import torch
import torch.nn as nn

from ... import ops
from ..utils.index import pd_indexing  # noqa: F401  (re-exported like the reference module)

f32 = torch.float32


def effective_weight(m):
    """Weight of a holder module whether or not weight-norm is applied: g * v / ||v|| (norm over all dims but 0)."""
    if hasattr(m, "weight_g"):
        return torch._weight_norm(m.weight_v, m.weight_g, 0).detach()
    return m.weight.detach()


class Conv1d(nn.Conv1d):
    """Parameter holder, kaiming-normal(relu) weights and zero bias (residual_block.py:27-38)."""

    def reset_parameters(self):
        nn.init.kaiming_normal_(self.weight, nonlinearity="relu")
        if self.bias is not None:
            nn.init.constant_(self.bias, 0.0)


class Conv1d1x1(Conv1d):
    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__(in_channels, out_channels, kernel_size=1, padding=0, dilation=1, bias=bias)


class Conv2d(nn.Conv2d):
    def reset_parameters(self):
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")
        if self.bias is not None:
            nn.init.constant_(self.bias, 0.0)


class Conv2d1x1(Conv2d):
    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__(in_channels, out_channels, kernel_size=1, padding=0, dilation=1, bias=bias)


def _gated_tail(block, y, c, residual):
    """+ aux 1x1, tanh * sigmoid, output 1x1 with (.. + residual) * sqrt(1/2) fused (residual_block.py:139-157)."""
    if c is not None:
        assert block.conv1x1_aux is not None
        ops.conv1d_f32(c, effective_weight(block.conv1x1_aux), None, out=y, accumulate=True)
    z = ops.gated_act_f32(y, ops.GATE_TANH_SIGMOID)
    out = block.conv1x1_out
    return ops.conv1d_f32(z, effective_weight(out), out.bias, residual=residual, out_scale=math.sqrt(0.5))


class FixedBlock(nn.Module):
    def __init__(self, residual_channels=64, gate_channels=128, skip_channels=64, aux_channels=80, kernel_size=3,
                 dilation=1, bias=True):
        super().__init__()
        padding = (kernel_size - 1) // 2 * dilation
        self.kernel_size, self.dilation = kernel_size, dilation
        self.conv = Conv1d(residual_channels, gate_channels, kernel_size, padding=padding, padding_mode="reflect",
                           dilation=dilation, bias=bias)
        self.conv1x1_aux = Conv1d1x1(aux_channels, gate_channels, bias=False) if aux_channels > 0 else None
        gate_out_channels = gate_channels // 2
        self.conv1x1_out = Conv1d1x1(gate_out_channels, residual_channels, bias=bias)
        self.conv1x1_skip = Conv1d1x1(gate_out_channels, skip_channels, bias=bias)

    def run(self, x, c):
        y = ops.conv1d_f32(x, effective_weight(self.conv), self.conv.bias, dilation=self.dilation,
                           tap_origin=(self.kernel_size - 1) // 2, pad_mode=ops.PAD_REFLECT)
        return _gated_tail(self, y, c, x)

    def forward(self, x, c):
        """(x, s) like residual_block.py:123-157; ``s`` (the dead skip branch) is evaluated only here, for callers
        that use a block stand-alone."""
        x = x.to(f32).contiguous()
        c = None if c is None else c.to(f32).contiguous()
        y = ops.conv1d_f32(x, effective_weight(self.conv), self.conv.bias, dilation=self.dilation,
                           tap_origin=(self.kernel_size - 1) // 2, pad_mode=ops.PAD_REFLECT)
        return _standalone_tail(self, y, c, x)


class AdaptiveBlock(nn.Module):
    def __init__(self, residual_channels=64, gate_channels=128, skip_channels=64, aux_channels=80, bias=True):
        super().__init__()
        self.convP = Conv1d1x1(residual_channels, gate_channels, bias=bias)  # past
        self.convC = Conv1d1x1(residual_channels, gate_channels, bias=bias)  # current
        self.convF = Conv1d1x1(residual_channels, gate_channels, bias=bias)  # future
        self.conv1x1_aux = Conv1d1x1(aux_channels, gate_channels, bias=False) if aux_channels > 0 else None
        gate_out_channels = gate_channels // 2
        self.conv1x1_out = Conv1d1x1(gate_out_channels, residual_channels, bias=bias)
        self.conv1x1_skip = Conv1d1x1(gate_out_channels, skip_channels, bias=bias)

    def stacked_taps(self):
        """[G, C, 3] weight (past, current, future) and the summed bias: one INDEXED conv instead of three 1x1s."""
        w = torch.stack([effective_weight(self.convP)[:, :, 0], effective_weight(self.convC)[:, :, 0],
                         effective_weight(self.convF)[:, :, 0]], dim=2).contiguous()
        b = None
        if self.convC.bias is not None:
            b = (self.convP.bias + self.convC.bias + self.convF.bias).detach()
        return w, b

    def run(self, x, c, idx):
        w, b = self.stacked_taps()
        y = ops.conv1d_f32(x, w, b, pad_mode=ops.PAD_INDEXED, idx=idx)
        return _gated_tail(self, y, c, x)

    def forward(self, xC, xP, xF, c):
        """(x, s) from explicit past/future tensors like residual_block.py:198-234."""
        xC, xP, xF = (t.to(f32).contiguous() for t in (xC, xP, xF))
        c = None if c is None else c.to(f32).contiguous()
        y = ops.conv1d_f32(xC, effective_weight(self.convC), self.convC.bias)
        ops.conv1d_f32(xP, effective_weight(self.convP), self.convP.bias, out=y, accumulate=True)
        ops.conv1d_f32(xF, effective_weight(self.convF), self.convF.bias, out=y, accumulate=True)
        return _standalone_tail(self, y, c, xC)


def _standalone_tail(block, y, c, residual):
    if c is not None:
        ops.conv1d_f32(c, effective_weight(block.conv1x1_aux), None, out=y, accumulate=True)
    z = ops.gated_act_f32(y, ops.GATE_TANH_SIGMOID)
    s = ops.conv1d_f32(z, effective_weight(block.conv1x1_skip), block.conv1x1_skip.bias)
    x = ops.conv1d_f32(z, effective_weight(block.conv1x1_out), block.conv1x1_out.bias, residual=residual,
                       out_scale=math.sqrt(0.5))
    return x, s


class ResidualBlocks(nn.Module):
    def __init__(self, blockA, cycleA, blockF, cycleF, cascade_mode=0, residual_channels=64, gate_channels=128,
                 skip_channels=64, aux_channels=80):
        super().__init__()
        cycleA, cycleF = max(cycleA, 1), max(cycleF, 1)
        assert blockA % cycleA == 0
        assert blockF % cycleF == 0
        self.blockA_per_cycle = blockA // cycleA
        blockF_per_cycle = blockF // cycleF
        adaptive = [AdaptiveBlock(residual_channels, gate_channels, skip_channels, aux_channels) for _ in range(blockA)]
        fixed = [FixedBlock(residual_channels, gate_channels, skip_channels, aux_channels,
                            dilation=2 ** (n % blockF_per_cycle)) for n in range(blockF)]
        if cascade_mode == 0:    # adaptive -> fixed
            self.conv_dilated = nn.ModuleList(adaptive + fixed)
            self.block_modes = [True] * blockA + [False] * blockF
        elif cascade_mode == 1:  # fixed -> adaptive
            self.conv_dilated = nn.ModuleList(fixed + adaptive)
            self.block_modes = [False] * blockF + [True] * blockA
        else:
            raise ValueError(f"Cascaded mode {cascade_mode} is not supported!")

    # ------------------------------------------------------------------ bf16 tensor-core path (NTC layout)
    def supports_bf16(self):
        blocks = list(self.conv_dilated)
        if not blocks:
            return True
        b0 = blocks[0]
        res = b0.conv1x1_out.out_channels
        gate = b0.conv1x1_out.in_channels * 2
        aux = b0.conv1x1_aux.in_channels if b0.conv1x1_aux is not None else 0
        fixed_ok = all(isinstance(b, AdaptiveBlock) or b.kernel_size == 3 for b in blocks)
        return res == 64 and gate == 128 and 0 < aux <= 320 and fixed_ok

    def _bf16_plan(self, aux_pad):
        """Packed bf16 weights per block.  ``aux_pad`` = padded aux width of the sample-rate path, or 0 = the frame-rate
        aux projection (w1p holds the taps only; the aux weights go into ``aux_weights_bf16``)."""
        key = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (aux_pad,)
        if getattr(self, "_plan_key", None) == key:
            return self._plan
        plan = []
        with torch.no_grad():
            for block, adaptive in zip(self.conv_dilated, self.block_modes):
                if adaptive:
                    w_taps, b1 = block.stacked_taps()
                else:
                    w_taps, b1 = effective_weight(block.conv).contiguous(), block.conv.bias
                w_aux = None
                if aux_pad:
                    w_aux = effective_weight(block.conv1x1_aux)[:, :, 0]
                    if aux_pad > w_aux.shape[1]:
                        w_aux = torch.nn.functional.pad(w_aux, (0, aux_pad - w_aux.shape[1]))
                    w_aux = w_aux.to(f32).contiguous()
                w_out = effective_weight(block.conv1x1_out)[:, :, 0]
                w1p, woutp = ops.usfgan_pack_block(w_taps.to(f32), w_aux, w_out.to(f32).contiguous())
                zeros = torch.zeros(w_taps.shape[0], device=w_taps.device, dtype=f32)
                plan.append(dict(w1p=w1p, woutp=woutp,
                                 bias1=(b1.detach().to(f32).contiguous() if b1 is not None else zeros),
                                 bout=block.conv1x1_out.bias.detach().to(f32).contiguous(),
                                 dilation=getattr(block, "dilation", 1), adaptive=adaptive))
        self._plan, self._plan_key = plan, key
        return plan

    def aux_weights(self, aux_pad):
        """[128 * blocks, aux_pad] fp32: the blocks' conv1x1_aux weights stacked (rows of the frame-rate projection)."""
        rows = []
        for block in self.conv_dilated:
            w = effective_weight(block.conv1x1_aux)[:, :, 0].to(f32)
            rows.append(torch.nn.functional.pad(w, (0, aux_pad - w.shape[1])) if aux_pad > w.shape[1] else w)
        return torch.cat(rows, dim=0)

    def forward_ntc_bf16(self, xb, auxb, d, idx_cache=None, relu_last=False, frames=None, frames_block0=0):
        """xb [B,T,64] bf16, auxb [B,T,A8] bf16 (A rounded up to a multiple of 8) or ``frames`` (ops.UsfganAuxFrames, this
        stack's first block at index ``frames_block0``), d (B,1,T) fp32 -> [B,T,64] bf16.
        One svsk_usfgan_block_bf16 launch per block; activations ping-pong between two buffers."""
        plan = self._bf16_plan(0 if frames is not None else auxb.shape[2])
        idx_cache = {} if idx_cache is None else idx_cache
        cur, nxt = xb, torch.empty_like(xb)
        a_idx = 0
        for n_blk, pw in enumerate(plan):
            idx = None
            if pw["adaptive"]:
                dil = 2 ** (a_idx % self.blockA_per_cycle)
                if dil not in idx_cache:
                    idx_cache[dil] = ops.pd_index(d.to(f32).contiguous(), dil)
                idx = idx_cache[dil]
                a_idx += 1
            ops.usfgan_block_bf16(cur, nxt, auxb, pw["w1p"], pw["woutp"], pw["bias1"], pw["bout"],
                                  dilation=pw["dilation"], idx=idx, out_relu=(relu_last and n_blk == len(plan) - 1),
                                  frames=frames, frames_block=frames_block0 + n_blk)
            cur, nxt = nxt, cur
        return cur

    def forward(self, x, c, d, batch_index=None, ch_index=None, idx_cache=None):
        """x (B,C,T), c (B,aux,T), d (B,1,T) -> (B,C,T)   (residual_block.py:311-336).
        ``idx_cache``: dict dilation -> (idx_past, idx_future), shared across stacks of one generator call."""
        x = x.to(f32).contiguous()
        c = c.to(f32).contiguous()
        idx_cache = {} if idx_cache is None else idx_cache
        a_idx = 0
        for block, adaptive in zip(self.conv_dilated, self.block_modes):
            if adaptive:
                dil = 2 ** (a_idx % self.blockA_per_cycle)
                if dil not in idx_cache:
                    idx_cache[dil] = ops.pd_index(d.to(f32).contiguous(), dil)
                x = block.run(x, c, idx_cache[dil])
                a_idx += 1
            else:
                x = block.run(x, c)
        return x


class PeriodicityEstimator(nn.Module):
    def __init__(self, in_channels, residual_channels=64, conv_layers=3, kernel_size=5, dilation=1,
                 padding_mode="replicate"):
        super().__init__()
        self.kernel_size, self.dilation, self.padding_mode, self.conv_layers = kernel_size, dilation, padding_mode, conv_layers
        modules = []
        for n in range(conv_layers):
            conv = Conv1d(in_channels, residual_channels, kernel_size=kernel_size, dilation=dilation,
                          padding=kernel_size // 2 * dilation, padding_mode=padding_mode)
            if n != conv_layers - 1:
                act = nn.ReLU(inplace=True)
            else:
                nn.init.normal_(conv.weight, std=1e-4)  # sigmoid(~0) = 0.5 at init (residual_block.py:378-382)
                act = nn.Sigmoid()
            modules += [conv, act]
            in_channels = residual_channels
        self.layers = nn.Sequential(*modules)

    def supports_bf16(self):
        convs = [self.layers[2 * n] for n in range(self.conv_layers)]
        return all(c.out_channels % 16 == 0 and c.out_channels <= 256 and c.kernel_size[0] <= 8 for c in convs)

    def forward_ntc_bf16(self, auxb):
        """auxb [B,T,A8] bf16 (A padded to a multiple of 8 with zeros) -> a [B,T,Cout] bf16, tcgen05 conv kernel."""
        key = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (auxb.shape[2],)
        if getattr(self, "_plan_key", None) != key:
            plan = []
            with torch.no_grad():
                cin = auxb.shape[2]
                for n in range(self.conv_layers):
                    conv = self.layers[2 * n]
                    w = effective_weight(conv).to(f32)
                    if cin > w.shape[1]:
                        w = torch.nn.functional.pad(w, (0, 0, 0, cin - w.shape[1]))
                    plan.append((ops.conv1d_pack_bf16(w.contiguous()), conv.bias.detach().to(f32).contiguous(),
                                 conv.out_channels))
                    cin = conv.out_channels
            self._plan, self._plan_key = plan, key
        pad = {"replicate": ops.PAD_REPLICATE, "reflect": ops.PAD_REFLECT, "zeros": ops.PAD_ZEROS}[self.padding_mode]
        x = auxb
        for n, (wp, bias, cout) in enumerate(self._plan):
            act = ops.ACT_SIGMOID if n == self.conv_layers - 1 else ops.ACT_RELU
            x = ops.conv1d_bf16(x, wp, bias, cout, self.kernel_size, dilation=self.dilation,
                                tap_origin=self.kernel_size // 2, pad_mode=pad, act=act)
        return x

    def forward(self, x):
        pad = {"replicate": ops.PAD_REPLICATE, "reflect": ops.PAD_REFLECT, "zeros": ops.PAD_ZEROS}[self.padding_mode]
        x = x.to(f32).contiguous()
        for n in range(self.conv_layers):
            conv = self.layers[2 * n]
            act = ops.ACT_SIGMOID if n == self.conv_layers - 1 else ops.ACT_RELU
            x = ops.conv1d_f32(x, effective_weight(conv), conv.bias, dilation=self.dilation,
                               tap_origin=self.kernel_size // 2, pad_mode=pad, act=act)
        return x
