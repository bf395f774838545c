from .generator import CascadeHnUSFGANGenerator, ParallelHnUSFGANGenerator, USFGANGenerator  # noqa: F401
