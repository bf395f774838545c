"""B200-native gated dilated-Conv1d stacks for sarulab-speech/ensemble_svs_with_interactions.

Drop-in replacements (same constructor kwargs, forward/inference signatures and state_dict layout) for
``nnsvs.diffsinger.{DiffNet,GaussianDiffusion,MultiSpeakerGaussianDiffusion}``, ``nnsvs.wavenet.WaveNet`` and
``nnsvs.usfgan.{USFGANWrapper, models.*Generator}``; the compute runs in hand-written sm_100a CUDA kernels behind the
C ABI in include/svsk.h (libsvsk.so).  No CPU path, no PyTorch-op fallback.
"""
from . import _lib  # noqa: F401
from .base import BaseModel, PredictionType  # noqa: F401

__all__ = ["BaseModel", "PredictionType"]
