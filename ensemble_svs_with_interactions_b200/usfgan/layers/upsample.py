"""Aux-feature upsampler with the module tree of ``nnsvs.usfgan.layers.upsample`` (upsample.py:15-194).

``ConvInUpsampleNetwork``: a k = 2*window+1 Conv1d without padding or bias at frame rate, then for every scale ``s`` a
nearest-neighbour stretch followed by a single-channel (1, 2s+1) smoothing filter shared by all aux channels.  The
sub-modules hold the parameters (``conv_in``, ``upsample.up_layers.{1,3,..}``); the arithmetic is svsk_conv1d_f32
(VALID mode) + svsk_upsample_smooth_f32 (stretch and smoothing fused, no stretched intermediate).
"""
import numpy as np
import torch

from ... import ops
from .residual_block import Conv1d, effective_weight


class Stretch2d(torch.nn.Module):
    """Parameter-free placeholder keeping ``up_layers`` indices identical to upsample.py:84-101."""

    def __init__(self, x_scale, y_scale, mode="nearest"):
        super().__init__()
        self.x_scale, self.y_scale, self.mode = x_scale, y_scale, mode


class Conv2d(torch.nn.Conv2d):
    """Smoothing-filter holder; box-filter init 1/prod(kernel) as upsample.py:54-58."""

    def reset_parameters(self):
        self.weight.data.fill_(1.0 / np.prod(self.kernel_size))
        if self.bias is not None:
            torch.nn.init.constant_(self.bias, 0.0)


class UpsampleNetwork(torch.nn.Module):
    def __init__(self, upsample_scales, nonlinear_activation=None, nonlinear_activation_params={},
                 interpolate_mode="nearest", freq_axis_kernel_size=1, use_causal_conv=False):
        super().__init__()
        if nonlinear_activation is not None or interpolate_mode != "nearest" or freq_axis_kernel_size != 1 \
                or use_causal_conv:
            raise NotImplementedError("libsvsk upsampler supports the recipes' configuration only: nearest stretch, "
                                      "no nonlinearity, freq kernel 1, non-causal")
        self.upsample_scales = list(upsample_scales)
        self.up_layers = torch.nn.ModuleList()
        for scale in self.upsample_scales:
            self.up_layers += [Stretch2d(scale, 1, interpolate_mode)]
            self.up_layers += [Conv2d(1, 1, kernel_size=(1, scale * 2 + 1), padding=(0, scale), bias=False)]

    def forward(self, c):
        c = c.to(torch.float32).contiguous()
        for n, scale in enumerate(self.upsample_scales):
            taps = effective_weight(self.up_layers[2 * n + 1]).reshape(-1).contiguous()
            c = ops.upsample_smooth_f32(c, taps, scale)
        return c

    def supports_fused(self, channels):
        return ops.upsample_fused_supported(self.upsample_scales, channels)

    def forward_ntc(self, c, want_bf16=True, want_f32=False):
        """All stages in one pass, channel-last output (svsk_upsample_fused): c (B, C, T') fp32 ->
        (B, T' * prod(scales), C rounded up to 8) bf16 and/or fp32."""
        taps = torch.cat([effective_weight(self.up_layers[2 * n + 1]).reshape(-1) for n in range(len(self.upsample_scales))])
        return ops.upsample_fused(c.to(torch.float32).contiguous(), taps.to(torch.float32).contiguous(), self.upsample_scales,
                                  want_bf16=want_bf16, want_f32=want_f32)


class ConvInUpsampleNetwork(torch.nn.Module):
    def __init__(self, upsample_scales, nonlinear_activation=None, nonlinear_activation_params={},
                 interpolate_mode="nearest", freq_axis_kernel_size=1, aux_channels=80, aux_context_window=0,
                 use_causal_conv=False):
        super().__init__()
        self.aux_context_window = aux_context_window
        self.use_causal_conv = use_causal_conv and aux_context_window > 0
        if self.use_causal_conv:
            raise NotImplementedError("causal aux upsampling is not used by the recipes and not built")
        self.conv_in = Conv1d(aux_channels, aux_channels, kernel_size=2 * aux_context_window + 1, bias=False)
        self.upsample = UpsampleNetwork(upsample_scales, nonlinear_activation, nonlinear_activation_params,
                                        interpolate_mode, freq_axis_kernel_size, use_causal_conv)

    def conv_in_frames(self, c):
        """c (B, C, T' + 2*window) -> conv_in's output (B, C, T') at frame rate (upsample.py:186-187)."""
        return ops.conv1d_f32(c.to(torch.float32).contiguous(), effective_weight(self.conv_in), None,
                              pad_mode=ops.PAD_VALID)

    def forward(self, c):
        """c (B, C, T' + 2*window) -> (B, C, T' * prod(scales))."""
        return self.upsample(self.conv_in_frames(c))

    def supports_fused(self):
        return self.upsample.supports_fused(self.conv_in.out_channels)

    def forward_ntc_bf16(self, c, cin=None):
        """c (B, C, T' + 2*window) -> (B, T' * prod(scales), C rounded up to 8) bf16, no fp32 sample-rate intermediate.
        ``cin`` = conv_in_frames(c) if the caller already has it."""
        return self.upsample.forward_ntc(self.conv_in_frames(c) if cin is None else cin)[0]
