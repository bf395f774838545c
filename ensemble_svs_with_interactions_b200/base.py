"""Model interface the drop-in classes satisfy — mirror of nnsvs/base.py:11-157 (PredictionType, BaseModel)."""
from enum import Enum

from torch import nn


class PredictionType(Enum):
    """nnsvs/base.py:11-76."""

    DETERMINISTIC = 1
    PROBABILISTIC = 2
    MULTISTREAM_HYBRID = 3
    DIFFUSION = 4


class BaseModel(nn.Module):
    """nnsvs/base.py:79-157: forward / inference / preprocess_target / prediction_type / ... defaults."""

    def forward(self, x, lengths=None, y=None):
        pass

    def inference(self, x, lengths=None):
        return self(x, lengths)

    def preprocess_target(self, y):
        return y

    def prediction_type(self):
        return PredictionType.DETERMINISTIC

    def is_autoregressive(self):
        return False

    def has_residual_lf0_prediction(self):
        return False
