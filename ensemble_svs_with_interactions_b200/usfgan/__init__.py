"""``USFGANWrapper`` — drop-in for ``nnsvs.usfgan.USFGANWrapper`` (nnsvs/usfgan/__init__.py:7-65): the vocoder entry
``gen.predict_waveform`` calls (gen.py:1694)."""
import numpy as np
import torch
from torch import nn

from .. import ops
from .utils import SignalGenerator, dilated_factor


class USFGANWrapper(nn.Module):
    def __init__(self, config, generator):
        super().__init__()
        self.generator = generator
        self.config = config

    @torch.no_grad()
    def inference(self, f0, aux_feats, *, noise=None):
        """f0: numpy (T, 1); aux_feats: Tensor (T, C) on the generator's device -> waveform (1, 1, T * hop).
        ``noise`` (parity harness only): {"sine": ..., "noise": ...} tensors (1, 1, T * hop) that replace the source
        signal's two Gaussian draws."""
        data = self.config.data
        assert data.sine_f0_type in ["contf0", "cf0", "f0"]
        assert data.df_f0_type in ["contf0", "cf0", "f0"]
        if "aux_context_window" not in self.config.generator:
            raise NotImplementedError("SiFi-GAN generators are not part of this build (uSFGAN family only)")
        device = aux_feats.device
        window = self.config.generator.aux_context_window
        c = nn.functional.pad(aux_feats.unsqueeze(0).transpose(2, 1), (window, window), mode="replicate").to(device)
        if device.type == "cuda":
            # dilated_factor (features.py:56-75) on the device, in float64 from the caller's own values like the
            # reference's numpy expression (the tap offsets round(d * dilation) must come out identical)
            f64 = torch.as_tensor(np.asarray(f0).reshape(1, -1), dtype=torch.float64).to(device)
            _, df = ops.usfgan_source(f64, hop=data.hop_size, sample_rate=data.sample_rate, dense_factor=data.dense_factor,
                                      want_sine=False)
        else:
            df = dilated_factor(np.squeeze(f0.copy()), data.sample_rate, data.dense_factor).repeat(data.hop_size, axis=0)
            df = torch.FloatTensor(df).view(1, 1, -1).to(device)
        f0 = torch.FloatTensor(f0).unsqueeze(0).transpose(2, 1).to(device)
        signal_generator = SignalGenerator(sample_rate=data.sample_rate, hop_size=data.hop_size,
                                           sine_amp=data.sine_amp, noise_amp=data.noise_amp,
                                           signal_types=data.signal_types)
        signal_generator.injected_noise = noise
        in_signal = signal_generator(f0)
        if getattr(self.generator, "supports_wave_only", False):
            return self.generator(in_signal, c.contiguous(), df, wave_only=True)[0]
        return self.generator(in_signal, c.contiguous(), df)[0]

    @torch.no_grad()
    def inference_batch(self, f0, aux_feats, *, noise=None):
        """Batched form of ``inference`` (the reference wrapper has no batch dimension and the
        multi-track synthesis loops over tracks, synthesis_multitrack.py:113-118).

        f0: (B, T, 1) numpy array or tensor, aux_feats: (B, T, C) tensor on the generator's device; all tracks of a call
        have the same number of frames (pad and trim outside, as the reference pads).  Returns (B, 1, T * hop).
        Everything runs on the device: no numpy round trip for the dilation factors."""
        data = self.config.data
        assert data.sine_f0_type in ["contf0", "cf0", "f0"]
        assert data.df_f0_type in ["contf0", "cf0", "f0"]
        if "aux_context_window" not in self.config.generator:
            raise NotImplementedError("SiFi-GAN generators are not part of this build (uSFGAN family only)")
        device = aux_feats.device
        if aux_feats.dim() != 3:
            raise ValueError(f"aux_feats must be (B, T, C), got {tuple(aux_feats.shape)}")
        f0 = torch.as_tensor(f0, dtype=torch.float32).to(device)
        if f0.dim() != 3 or f0.shape[:2] != aux_feats.shape[:2] or f0.shape[2] != 1:
            raise ValueError(f"f0 must be (B, T, 1) matching aux_feats, got {tuple(f0.shape)}")
        window = self.config.generator.aux_context_window
        f0 = f0.transpose(2, 1).contiguous()                                      # (B, 1, T)
        # dilated_factor (features.py:56-75) on the device: unvoiced frames count as fs / dense_factor, i.e. factor 1
        # (in float64 like the reference's numpy expression, then rounded to fp32 once: the tap offsets round(d * dilation)
        # must come out identical)
        _, df = ops.usfgan_source(f0[:, 0].double().contiguous(), hop=data.hop_size, sample_rate=data.sample_rate,
                                  dense_factor=data.dense_factor, want_sine=False)
        c = nn.functional.pad(aux_feats.transpose(2, 1), (window, window), mode="replicate")
        signal_generator = SignalGenerator(sample_rate=data.sample_rate, hop_size=data.hop_size,
                                           sine_amp=data.sine_amp, noise_amp=data.noise_amp,
                                           signal_types=data.signal_types)
        signal_generator.injected_noise = noise
        in_signal = signal_generator(f0)
        if getattr(self.generator, "supports_wave_only", False):
            return self.generator(in_signal, c.contiguous(), df.contiguous(), wave_only=True)[0]
        return self.generator(in_signal, c.contiguous(), df.contiguous())[0]
