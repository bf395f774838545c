"""TEST INFRASTRUCTURE ONLY — loader for the *real* reference modules.

The reference (`/root/reference`, pure Python/PyTorch) cannot be imported with a
plain ``import nnsvs``: ``nnsvs/__init__.py:1`` pulls in ``nnsvs.util`` which needs
pyworld / hydra / omegaconf, and ``nnsvs/usfgan/utils/utils.py:15`` needs h5py,
``nnsvs/usfgan/models/discriminator.py:16`` needs tkinter.  None of those are on
the hot path, so this shim registers a namespace stub for ``nnsvs`` (skipping its
``__init__``) plus empty stubs for ``h5py``/``tkinter``.

Where the reference comes from, in order: ``$SVSK_REFERENCE_ROOT``; ``baseline/_ref`` (the unmodified reference
installed by ``pip install --no-index --no-deps --target baseline/_ref`` — done by ``__graft_entry__.build()`` in the
build container; git-ignored, but it travels to the GPU box with the snapshot); ``/root/reference`` (build container
only).

Users:

* ``oracle/make_golden.py`` — generates ``tests/golden/*.npz`` from the reference.
* ``tests/test_oracle_vs_reference.py`` — pins the oracle restatement against the
  live reference (skipped when no reference tree is found).
* ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` / ``reference_gpu_eager`` legs — time the
  reference's own modules (``kind: "reference"``); the oracle port is the fallback when no tree is found.

Nothing in the product package may import this file.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    cands = [os.environ.get("SVSK_REFERENCE_ROOT"), os.path.join(_REPO, "baseline", "_ref"), "/root/reference"]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "nnsvs", "diffsinger")):
            return c
    return cands[-1]


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "nnsvs", "diffsinger"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference():
    """Returns a namespace with the reference hot-path classes (unmodified code)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    if "nnsvs" not in sys.modules or not getattr(sys.modules["nnsvs"], "_svsk_shim", False):
        pkg = types.ModuleType("nnsvs")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "nnsvs")]
        pkg._svsk_shim = True
        sys.modules["nnsvs"] = pkg
    _stub("h5py")
    _stub("tkinter", W="w")
    # nnsvs/model.py:9,12 import split_streams (needs nnmnkwii) and init_weights (nnsvs.util needs pyworld / hydra);
    # neither is on the FFConvLSTM path when init_type == "none" (util.py:40-41 returns at once).
    _stub("nnsvs.multistream", split_streams=None)

    def _init_weights(net, init_type="none", init_gain=0.02):
        assert init_type == "none", "the shim only covers init_type='none'"
    _stub("nnsvs.util", init_weights=_init_weights)

    ns = types.SimpleNamespace()
    from nnsvs.diffsinger.denoiser import DiffNet
    from nnsvs.diffsinger.diffusion import GaussianDiffusion
    from nnsvs.wavenet import WaveNet
    import nnsvs.usfgan.models.generator as gen
    import nnsvs.usfgan.layers.residual_block as rb
    import nnsvs.usfgan.layers.upsample as up
    import nnsvs.usfgan.utils.index as index
    import nnsvs.usfgan.utils.features as features
    from nnsvs.usfgan import USFGANWrapper
    from nnsvs.model import FFConvLSTM
    from nnsvs.dsp import lowpass_filter
    from nnsvs.postfilters import variance_scaling

    # pd_indexing/index_initial call .cuda() whenever CUDA is visible
    # (nnsvs/usfgan/utils/index.py:32-33,44-45,81-83); the CPU oracle use of the
    # reference therefore needs CUDA hidden.  We do not patch the reference; callers
    # run this with CUDA_VISIBLE_DEVICES="" when a GPU is present.
    ns.DiffNet = DiffNet
    ns.GaussianDiffusion = GaussianDiffusion
    ns.WaveNet = WaveNet
    ns.USFGANGenerator = gen.USFGANGenerator
    ns.CascadeHnUSFGANGenerator = gen.CascadeHnUSFGANGenerator
    ns.ParallelHnUSFGANGenerator = gen.ParallelHnUSFGANGenerator
    ns.FixedBlock = rb.FixedBlock
    ns.AdaptiveBlock = rb.AdaptiveBlock
    ns.ResidualBlocks = rb.ResidualBlocks
    ns.PeriodicityEstimator = rb.PeriodicityEstimator
    ns.ConvInUpsampleNetwork = up.ConvInUpsampleNetwork
    ns.pd_indexing = index.pd_indexing
    ns.index_initial = index.index_initial
    ns.SignalGenerator = features.SignalGenerator
    ns.dilated_factor = features.dilated_factor
    ns.USFGANWrapper = USFGANWrapper
    ns.FFConvLSTM = FFConvLSTM
    ns.lowpass_filter = lowpass_filter
    ns.variance_scaling = variance_scaling
    return ns
