"""``USFGANWrapper`` — drop-in for ``nnsvs.usfgan.USFGANWrapper`` (nnsvs/usfgan/__init__.py:7-65): the vocoder entry
``gen.predict_waveform`` calls (gen.py:1694)."""
import numpy as np
import torch
from torch import nn

from .utils import SignalGenerator, dilated_factor


class USFGANWrapper(nn.Module):
    def __init__(self, config, generator):
        super().__init__()
        self.generator = generator
        self.config = config

    @torch.no_grad()
    def inference(self, f0, aux_feats):
        """f0: numpy (T, 1); aux_feats: Tensor (T, C) on the generator's device -> waveform (1, 1, T * hop)."""
        data = self.config.data
        assert data.sine_f0_type in ["contf0", "cf0", "f0"]
        assert data.df_f0_type in ["contf0", "cf0", "f0"]
        if "aux_context_window" not in self.config.generator:
            raise NotImplementedError("SiFi-GAN generators are not part of this build (uSFGAN family only)")
        device = aux_feats.device
        window = self.config.generator.aux_context_window
        df = dilated_factor(np.squeeze(f0.copy()), data.sample_rate, data.dense_factor).repeat(data.hop_size, axis=0)
        c = nn.functional.pad(aux_feats.unsqueeze(0).transpose(2, 1), (window, window), mode="replicate").to(device)
        df = torch.FloatTensor(df).view(1, 1, -1).to(device)
        f0 = torch.FloatTensor(f0).unsqueeze(0).transpose(2, 1).to(device)
        signal_generator = SignalGenerator(sample_rate=data.sample_rate, hop_size=data.hop_size,
                                           sine_amp=data.sine_amp, noise_amp=data.noise_amp,
                                           signal_types=data.signal_types)
        in_signal = signal_generator(f0)
        if getattr(self.generator, "supports_wave_only", False):
            return self.generator(in_signal, c.contiguous(), df, wave_only=True)[0]
        return self.generator(in_signal, c.contiguous(), df)[0]
