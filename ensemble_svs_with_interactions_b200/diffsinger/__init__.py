from .denoiser import DiffNet
from .diffusion import GaussianDiffusion, MultiSpeakerGaussianDiffusion

__all__ = ["DiffNet", "GaussianDiffusion", "MultiSpeakerGaussianDiffusion"]
