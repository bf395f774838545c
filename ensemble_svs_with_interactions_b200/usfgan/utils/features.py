"""uSFGAN front-end helpers with the API of ``nnsvs.usfgan.utils.features``
(``dilated_factor`` features.py:56-75, ``SignalGenerator`` features.py:78-180).

On a CUDA device the sine source and the dilation factors are one libsvsk kernel pair (``svsk_usfgan_source``,
SURVEY.md §8(f) row 2): a per-track fp64 scan over the frames and one pass over the samples — the phase is the fp64
prefix sum rounded to fp32 per sample, which is bit for bit what ``torch.cumsum`` gives the reference on a CPU host
(CUDA's own fp32 parallel scan drifts from it).  The numpy ``dilated_factor`` and the torch expressions below remain for
host-side callers (numpy inputs, CPU tensors in the host-logic tests).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from ... import ops


def dilated_factor(batch_f0, fs, dense_factor):
    """Dilation in samples that puts ``dense_factor`` taps in one pitch period: fs / (f0 * dense_factor).
    Unvoiced frames (f0 == 0) are first set to fs / dense_factor (factor 1), IN PLACE as the reference does."""
    unvoiced = batch_f0 == 0
    batch_f0[unvoiced] = fs / dense_factor
    factors = np.full(batch_f0.shape, float(fs)) / batch_f0 / dense_factor
    if not np.all(factors > 0):
        raise AssertionError("dilated factors must be positive")
    return factors


class SignalGenerator:
    """Builds the generator's input channels from frame-level F0 (B, 1, frames): NSF-style sine, Gaussian noise
    and/or a V/UV mask, each upsampled by ``hop_size`` with nearest-neighbour hold, concatenated on dim 1."""

    def __init__(self, sample_rate=24000, hop_size=120, sine_amp=0.1, noise_amp=0.003, signal_types=["sine", "noise"]):
        self.sample_rate = sample_rate
        self.hop_size = hop_size
        self.sine_amp = sine_amp
        self.noise_amp = noise_amp
        self.signal_types = signal_types
        # parity harness only: {"sine": (B,1,T) tensor, "noise": (B,1,T) tensor} replaces the two Gaussian draws
        # (the reference draws them from torch's global generator, features.py:131,160 — CPU and CUDA streams differ)
        self.injected_noise = None

    def _randn(self, which, shape, device):
        if self.injected_noise is not None and which in self.injected_noise:
            z = self.injected_noise[which].to(device=device, dtype=torch.float32)
            if tuple(z.shape) != tuple(shape):
                raise ValueError(f"injected {which} noise has shape {tuple(z.shape)}, expected {tuple(shape)}")
            return z
        return torch.randn(shape, device=device)

    def _hold(self, frames):
        # F.interpolate(nearest) rather than repeat_interleave: identical index arithmetic to the reference
        return F.interpolate(frames, frames.shape[-1] * self.hop_size)

    def _voiced_mask(self, f0):
        return self._hold((f0 > 0).to(f0.dtype))

    @torch.no_grad()
    def random_noise(self, f0):
        return self._randn("noise", (f0.shape[0], 1, f0.shape[-1] * self.hop_size), f0.device)

    @torch.no_grad()
    def vuv_binary(self, f0):
        return self._voiced_mask(f0)

    @torch.no_grad()
    def sinusoid(self, f0):
        if f0.is_cuda:
            B, _, Fr = f0.shape
            noise = self._randn("sine", (B, 1, Fr * self.hop_size), f0.device) if self.noise_amp > 0 else None
            sine, _ = ops.usfgan_source(f0[:, 0].double().contiguous(), hop=self.hop_size, sample_rate=self.sample_rate,
                                        sine_amp=self.sine_amp, noise_amp=self.noise_amp, noise=noise, want_d=False)
            return sine
        voiced = self._voiced_mask(f0)
        cycles_per_sample = torch.remainder(self._hold(f0) / self.sample_rate, 1)
        phase = torch.cumsum(cycles_per_sample, dim=2) * 2 * math.pi
        wave = voiced * torch.sin(phase) * self.sine_amp
        if self.noise_amp > 0:
            # voiced frames get noise_amp, unvoiced ones a third of it (NSF)
            sigma = voiced * self.noise_amp + (1.0 - voiced) * self.noise_amp / 3.0
            wave = wave + self._randn("sine", wave.shape, f0.device) * sigma
        return wave

    @torch.no_grad()
    def __call__(self, f0):
        makers = {"noise": self.random_noise, "sine": self.sinusoid, "uv": self.vuv_binary}
        parts = [makers[kind](f0) for kind in self.signal_types if kind in makers]
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
