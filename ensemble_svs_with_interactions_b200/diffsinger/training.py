"""Training-time DiffNet forward with gradients.

The FORWARD pass always runs in libsvsk (same kernels as inference).  The BACKWARD pass is SURVEY.md §8(f) row 4
("training backward kernels", scheduled after rows a-e): until those dgrad/wgrad kernels exist, gradients are obtained
by re-evaluating the block stack with stock PyTorch ops inside ``backward`` and differentiating that.  Parameters are
passed through the autograd.Function, so DistributedDataParallel's per-parameter hooks fire and the NCCL gradient
all-reduce overlaps the backward pass exactly as with the reference model (train_util.py:1444-1446).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _mish(x):
    return x * torch.tanh(F.softplus(x))


def torch_restatement(net, spec, t, cond):
    """Differentiable re-statement of DiffNet.forward (used ONLY inside backward)."""
    C = net.residual_channels
    x = F.relu(F.conv1d(spec[:, 0], net.input_projection.weight, net.input_projection.bias))
    half = C // 2
    freq = torch.exp(torch.arange(half, device=spec.device) * -(math.log(10000) / (half - 1)))
    arg = t.to(torch.float32)[:, None] * freq[None, :]
    e = torch.cat((arg.sin(), arg.cos()), dim=-1)
    e = F.linear(_mish(F.linear(e, net.mlp[0].weight, net.mlp[0].bias)), net.mlp[2].weight, net.mlp[2].bias)
    skip = 0
    for layer in net.residual_layers:
        dp = F.linear(e, layer.diffusion_projection.weight, layer.diffusion_projection.bias)[:, :, None]
        y = F.conv1d(x + dp, layer.dilated_conv.weight, layer.dilated_conv.bias, padding=layer.dilation,
                     dilation=layer.dilation)
        y = y + F.conv1d(cond, layer.conditioner_projection.weight, layer.conditioner_projection.bias)
        z = torch.sigmoid(y[:, :C]) * torch.tanh(y[:, C:])
        o = F.conv1d(z, layer.output_projection.weight, layer.output_projection.bias)
        x = (x + o[:, :C]) / math.sqrt(2.0)
        skip = skip + o[:, C:]
    x = skip / math.sqrt(len(net.residual_layers))
    x = F.relu(F.conv1d(x, net.skip_projection.weight, net.skip_projection.bias))
    x = F.conv1d(x, net.output_projection.weight, net.output_projection.bias)
    return x[:, None]


class _DiffNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, spec, t, cond, *params):
        ctx.net = net
        ctx.save_for_backward(spec, t, cond)
        with torch.no_grad():
            return net._forward_no_grad(spec, t, cond)

    @staticmethod
    def backward(ctx, grad_out):
        net = ctx.net
        spec, t, cond = ctx.saved_tensors
        params = list(net.parameters())
        with torch.enable_grad():
            spec_ = spec.detach().requires_grad_(ctx.needs_input_grad[1])
            cond_ = cond.detach().requires_grad_(ctx.needs_input_grad[3])
            use_bf16 = net.resolved_precision() == "bf16"
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=use_bf16):
                out = torch_restatement(net, spec_, t, cond_)
            wanted = [x for x in (spec_, cond_) if x.requires_grad] + [p for p in params if p.requires_grad]
            grads = list(torch.autograd.grad(out, wanted, grad_out.to(out.dtype), allow_unused=True))
        g_spec = grads.pop(0) if spec_.requires_grad else None
        g_cond = grads.pop(0) if cond_.requires_grad else None
        g_params = [grads.pop(0) if p.requires_grad else None for p in params]
        return (None, g_spec, None, g_cond, *g_params)


def diffnet_forward_with_grad(net, spec, diffusion_step, cond):
    t = diffusion_step.reshape(-1).to(torch.int64)
    return _DiffNetFn.apply(net, spec, t, cond, *net.parameters())
