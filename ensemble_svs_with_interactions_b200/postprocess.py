"""Acoustic post-processing between the diffusion models and the vocoder, on the device and over whole batches.

SURVEY.md §8(f) row 3, second part: in the reference these run per utterance and per feature dimension in numpy / scipy on
the host, between the acoustic model and the vocoder (``gen.postprocess_acoustic``, gen.py:1394-1418 and 1500-1513):

* ``variance_scaling``  — ``nnsvs.postfilters.variance_scaling`` (postfilters.py:9-46), the GV post-filter;
* ``lowpass_filter``    — ``nnsvs.dsp.lowpass_filter`` (dsp.py:10-33), zero-phase Butterworth smoothing of every
  trajectory (``trajectory_smoothing``, cutoff 50 Hz for mgc / bap and 20 Hz for lf0 at 200 frames per second).

Same names and argument meaning; the arrays are CUDA tensors ``[B, T, D]`` (tracks x frames x dims) with optional
per-track ``lengths`` instead of one ``[T, D]`` numpy array per call.  The filter design (a handful of float64 numbers) is
host code; the trajectories never leave the GPU.  Not here: the pyworld aperiodicity round trip (gen.py:1639-1670) and
the learned / merlin post-filters.
"""
from __future__ import annotations

from functools import lru_cache
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

__all__ = ["variance_scaling", "lowpass_filter", "butter_lowpass", "lfilter_zi", "StandardScaler", "MinMaxScaler"]


@lru_cache(maxsize=64)
def butter_lowpass(N: int, Wn: float) -> Tuple[Tuple[float, ...], Tuple[float, ...]]:
    """Digital Butterworth low-pass of order N with cutoff Wn (1 = Nyquist): what ``scipy.signal.butter(N, Wn, "lowpass")``
    returns (dsp.py:25).  Analog prototype poles on the unit circle, pre-warped cutoff, bilinear transform at fs = 2."""
    if not (0.0 < Wn < 1.0):
        raise ValueError(f"cutoff must lie strictly between 0 and the Nyquist frequency, got Wn = {Wn}")
    k = np.arange(-N + 1, N, 2)
    poles = -np.exp(1j * np.pi * k / (2 * N))
    warped = 4.0 * np.tan(np.pi * Wn / 2.0)
    poles = warped * poles
    gain = warped ** N
    zd = (4.0 + poles) / (4.0 - poles)
    gain_d = gain * np.real(1.0 / np.prod(4.0 - poles))
    b = gain_d * np.poly(-np.ones(N))
    a = np.real(np.poly(zd))
    return tuple(float(v) for v in b), tuple(float(v) for v in a)


def lfilter_zi(b: Sequence[float], a: Sequence[float]) -> np.ndarray:
    """Steady-state direct-form-II-transposed state for a unit step (``scipy.signal.lfilter_zi``), a[0] == 1."""
    b, a = np.asarray(b, dtype=np.float64), np.asarray(a, dtype=np.float64)
    n = len(a) - 1
    comp = np.zeros((n, n))
    comp[0] = -a[1:]
    comp[1:, :-1] = np.eye(n - 1)
    return np.linalg.solve(np.eye(n) - comp.T, b[1:] - a[1:] * b[0])


def _lengths(lengths, B, device):
    if lengths is None:
        return None
    t = torch.as_tensor(lengths, dtype=torch.int32)
    if t.numel() != B:
        raise ValueError(f"lengths must have one entry per track ({B}), got {t.numel()}")
    return t.to(device)


def _check(x, what):
    if not x.is_cuda:
        raise RuntimeError(f"{what}: features must be a CUDA tensor (libsvsk has no CPU path)")
    if x.dim() != 3:
        raise ValueError(f"{what}: expected [B, T, D], got {tuple(x.shape)}")
    return x.detach().float().contiguous()


def lowpass_filter(x: torch.Tensor, fs: int, cutoff: float = 5, N: int = 5, lengths=None) -> torch.Tensor:
    """dsp.py:10-33 on every trajectory x[b, :lengths[b], d].  Trajectories of at most max(len(a), len(b)) * (N // 2 + 1)
    frames come back unchanged, like the reference's early return."""
    x = _check(x, "lowpass_filter")
    nyquist = fs // 2
    b, a = butter_lowpass(int(N), float(cutoff) / nyquist)
    ntaps = max(len(a), len(b))
    min_len = max(ntaps * (N // 2 + 1), 3 * ntaps)     # filtfilt itself needs more than 3 * ntaps samples
    return ops.filtfilt_f32(x, b, a, lfilter_zi(b, a), min_len=min_len, lengths=_lengths(lengths, x.shape[0], x.device))


def variance_scaling(gv: torch.Tensor, feats: torch.Tensor, offset: int = 2, note_mask: Optional[torch.Tensor] = None,
                     lengths=None) -> torch.Tensor:
    """postfilters.py:9-46 per track.  ``note_mask`` [B, T] (bool / uint8) plays the part of ``note_frame_indices``:
    statistics and scaling use the marked frames only; a track without marked frames is returned unchanged."""
    feats = _check(feats, "variance_scaling")
    B, T, D = feats.shape
    gv = torch.as_tensor(gv, dtype=torch.float32).to(feats.device).contiguous()
    if gv.numel() != D:
        raise ValueError(f"gv must have one entry per feature dimension ({D}), got {gv.numel()}")
    if note_mask is not None:
        if tuple(note_mask.shape) != (B, T):
            raise ValueError(f"note_mask must be [B, T] = {(B, T)}, got {tuple(note_mask.shape)}")
        note_mask = note_mask.to(feats.device).to(torch.uint8).contiguous()
    return ops.variance_scaling_f32(feats, gv, offset=offset, note_mask=note_mask, lengths=_lengths(lengths, B, feats.device))


def _vec(v, device, D, what):
    t = torch.as_tensor(np.asarray(v, dtype=np.float64) if not torch.is_tensor(v) else v).to(torch.float32).reshape(-1).to(device).contiguous()
    if t.numel() != D:
        raise ValueError(f"{what} must have one entry per feature ({D}), got {t.numel()}")
    return t


class _Scaler:
    def _apply(self, x, a, b, mode):
        if not x.is_cuda:
            raise RuntimeError("scaler: features must be a CUDA tensor (libsvsk has no CPU path)")
        x = x.detach().float().contiguous()
        D = x.shape[-1]
        return ops.scale_features_f32(x, _vec(a, x.device, D, "scale"), _vec(b, x.device, D, "offset"), mode)


class StandardScaler(_Scaler):
    """nnsvs/util.py:272-292 on CUDA tensors [..., D]: transform (x - mean_) / scale_, inverse_transform x * scale_ + mean_."""

    def __init__(self, mean, var, scale):
        self.mean_, self.var_, self.scale_ = mean, var, scale

    def transform(self, x):
        return self._apply(x, self.scale_, self.mean_, 1)

    def inverse_transform(self, x):
        return self._apply(x, self.scale_, self.mean_, 0)


class MinMaxScaler(_Scaler):
    """nnsvs/util.py:316-339: transform scale_ * x + min_, inverse_transform (x - min_) / scale_."""

    def __init__(self, min, scale, data_min=None, data_max=None, feature_range=(0, 1)):
        self.min_, self.scale_ = min, scale
        self.data_min_, self.data_max_, self.feature_range = data_min, data_max, feature_range

    def transform(self, x):
        return self._apply(x, self.scale_, self.min_, 0)

    def inverse_transform(self, x):
        return self._apply(x, self.scale_, self.min_, 1)
