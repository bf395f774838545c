from .residual_block import *  # noqa: F401,F403
from .residual_block import (AdaptiveBlock, Conv1d, Conv1d1x1, FixedBlock, PeriodicityEstimator,  # noqa: F401
                             ResidualBlocks, effective_weight)
from . import upsample  # noqa: F401
from .upsample import ConvInUpsampleNetwork, UpsampleNetwork  # noqa: F401
