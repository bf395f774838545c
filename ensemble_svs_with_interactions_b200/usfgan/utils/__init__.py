from .features import SignalGenerator, dilated_factor  # noqa: F401
from .index import index_initial, pd_indexing  # noqa: F401
