"""CPU-only: drop-in boundary (constructor kwargs, state_dict layout, schedule buffers) against the golden fixtures."""
import inspect
import warnings

import numpy as np
import pytest
import torch

from ensemble_svs_with_interactions_b200.base import BaseModel, PredictionType
from ensemble_svs_with_interactions_b200.diffsinger import DiffNet, GaussianDiffusion, MultiSpeakerGaussianDiffusion
from ensemble_svs_with_interactions_b200.model import FFConvLSTM, MultiSpeakerFFConvLSTM
from ensemble_svs_with_interactions_b200.usfgan.models import (CascadeHnUSFGANGenerator, ParallelHnUSFGANGenerator,
                                                              USFGANGenerator)
from ensemble_svs_with_interactions_b200.wavenet import WaveNet, receptive_field_size
from tests.golden_util import Golden

warnings.filterwarnings("ignore", category=FutureWarning)


def _same_layout(module, golden_sd):
    sd = module.state_dict()
    assert list(sd.keys()) == list(golden_sd.keys())          # names AND order
    for k in sd:
        assert tuple(sd[k].shape) == tuple(golden_sd[k].shape), k
    module.load_state_dict(golden_sd, strict=True)


def test_diffnet_state_dict_layout():
    g = Golden("diffnet_small")
    _same_layout(DiffNet(**g.cfg), g.sd)


def test_diffnet_ctor_signature_matches_reference():
    names = list(inspect.signature(DiffNet.__init__).parameters)[1:6]
    assert names == ["in_dim", "encoder_hidden_dim", "residual_layers", "residual_channels", "dilation_cycle_length"]
    m = DiffNet()
    assert (m.in_dim, len(m.residual_layers), m.residual_channels) == (80, 20, 256)
    assert [l.dilation for l in m.residual_layers[:5]] == [1, 2, 4, 8, 1]
    assert torch.count_nonzero(m.output_projection.weight) == 0   # zero init (denoiser.py:99)
    assert 15.0e6 < sum(p.numel() for p in m.parameters()) < 15.2e6   # DiffNet 15.08 M


def test_gaussian_diffusion_buffers_and_layout():
    g = Golden("diffusion_small")
    m = GaussianDiffusion(20, 12, DiffNet(**g.cfg["denoiser"]), K_step=g.cfg["K_step"])
    sd = m.state_dict()
    for k in list(g.sd)[:12]:
        assert torch.equal(sd[k], g.sd[k]), k       # float64 schedule maths -> fp32, bit exact
    _same_layout(m, g.sd)
    assert m.prediction_type() == PredictionType.DIFFUSION
    assert isinstance(m, BaseModel) and not m.is_autoregressive() and not m.has_residual_lf0_prediction()
    with pytest.raises(NotImplementedError):
        GaussianDiffusion(20, 12, DiffNet(**g.cfg["denoiser"]), pndm_speedup=4)
    cos = GaussianDiffusion(20, 12, DiffNet(**g.cfg["denoiser"]), K_step=10, schedule_type="cosine")
    assert cos.betas.shape == (10,) and float(cos.betas.max()) <= 0.999 + 1e-6


def test_multispeaker_signature():
    names = list(inspect.signature(MultiSpeakerGaussianDiffusion.__init__).parameters)
    assert names[1:5] == ["in_dim", "out_dim", "denoise_fn", "speaker_embedding"]


@pytest.mark.parametrize("name", ["wavenet_small", "wavenet_test_shape"])
def test_wavenet_state_dict_layout(name):
    g = Golden(name)
    m = WaveNet(**g.cfg)
    _same_layout(m, g.sd)
    m.remove_weight_norm_()
    assert "first_conv.weight" in m.state_dict() and "first_conv.weight_g" not in m.state_dict()
    assert receptive_field_size(30, 3, 2) == 3070


def _hn_kwargs(c):
    return dict(harmonic_network_params=c["harmonic"], noise_network_params=c["noise"],
                filter_network_params=c["filt"], periodicity_estimator_params=c["pe"], **c["common"])


def test_usfgan_state_dict_layouts_with_and_without_weight_norm():
    g = Golden("usfgan_parallel_hn_small")
    m = ParallelHnUSFGANGenerator(**_hn_kwargs(g.cfg))
    _same_layout(m, g.sd)
    m.remove_weight_norm()
    _same_layout(m, Golden("usfgan_parallel_hn_small_nowm").sd)
    m.apply_weight_norm()
    assert list(m.state_dict().keys()).count("conv_first_sine.weight_g") == 1
    g = Golden("usfgan_cascade_hn_small")
    _same_layout(CascadeHnUSFGANGenerator(**_hn_kwargs(g.cfg)), g.sd)
    g = Golden("usfgan_plain_small")
    _same_layout(USFGANGenerator(source_network_params=g.cfg["source"], filter_network_params=g.cfg["filt"],
                                 **g.cfg["common"]), g.sd)


def test_recipe_vocoder_key_counts():
    """756 keys with weight norm, 484 after remove_weight_norm for the recipe ParallelHn generator."""
    pe = {"conv_layers": 3, "kernel_size": 5, "dilation": 1, "padding_mode": "replicate"}
    m = ParallelHnUSFGANGenerator(periodicity_estimator_params=pe)
    assert len(m.state_dict()) == 756
    m.remove_weight_norm()
    assert len(m.state_dict()) == 484
    dil = [b.dilation for b in m.filter_network.conv_dilated]
    assert dil[:10] == [2 ** i for i in range(10)] and dil[10] == 1
    with pytest.raises(TypeError):       # reference quirk kept: the default key "conv_blocks" is rejected
        ParallelHnUSFGANGenerator()


def test_default_dicts_are_not_mutated():
    pe = {"conv_layers": 3, "kernel_size": 5, "dilation": 1, "padding_mode": "replicate"}
    hp = {"blockA": 2, "cycleA": 1, "blockF": 0, "cycleF": 0, "cascade_mode": 0}
    ParallelHnUSFGANGenerator(harmonic_network_params=hp, periodicity_estimator_params=pe)
    assert set(hp) == {"blockA", "cycleA", "blockF", "cycleF", "cascade_mode"}


def test_dilated_factor_and_signal_generator_cpu():
    from ensemble_svs_with_interactions_b200.usfgan.utils import SignalGenerator, dilated_factor
    g = Golden("usfgan_frontend")
    cfg = g.cfg
    f0 = g.inp["f0"].numpy()
    df = dilated_factor(np.squeeze(f0.copy()), cfg["sample_rate"], cfg["dense_factor"]).repeat(cfg["hop_size"])
    assert np.array_equal(df, g.out["df"].numpy())
    sg = SignalGenerator(sample_rate=cfg["sample_rate"], hop_size=cfg["hop_size"], sine_amp=cfg["sine_amp"],
                         noise_amp=cfg["noise_amp"], signal_types=["sine", "noise"])
    torch.manual_seed(53)
    sig = sg(torch.FloatTensor(f0).unsqueeze(0).transpose(2, 1))
    assert torch.allclose(sig, g.out["in_signal"], atol=1e-6)


def test_pipeline_plan_batches():
    """Host side of the batched (song, track) synthesis: ranks get items longest-first, batches
    respect the frame budget after padding, every item is scheduled exactly once."""
    from ensemble_svs_with_interactions_b200.pipeline import plan_batches
    lengths = [6000, 5990, 3000, 3001, 2999, 100, 6000, 42]
    seen = []
    for rank in range(2):
        plans = plan_batches(lengths, max_frames=12000, world_size=2, rank=rank, multiple=4)
        for p in plans:
            assert p.frames % 4 == 0 and p.frames >= max(lengths[i] for i in p.items)
            assert len(p.items) == 1 or len(p.items) * p.frames <= 12000
            seen += p.items
    assert sorted(seen) == list(range(len(lengths)))
    # one rank, huge budget: a single batch padded to the longest item
    (p,) = plan_batches(lengths, max_frames=10 ** 9)
    assert sorted(p.items) == list(range(len(lengths))) and p.frames == 6000
    # an item longer than the budget still runs, alone
    plans = plan_batches([50, 5000, 60], max_frames=1000)
    assert [sorted(p.items) for p in plans] == [[1], [0, 2]]
    import pytest
    with pytest.raises(ValueError):
        plan_batches(lengths, max_frames=0)


@pytest.mark.parametrize("name", ["ffconvlstm_embed", "ffconvlstm_test_shape"])
def test_ffconvlstm_state_dict_layout(name):
    """same keys, order and shapes as nnsvs.model.FFConvLSTM (model.py:801-890)."""
    g = Golden(name)
    m = FFConvLSTM(**g.cfg)
    _same_layout(m, g.sd)
    assert m.prediction_type() == PredictionType.DETERMINISTIC and not m.is_autoregressive()
    assert m.resolved_precision() == "fp32"          # widths 24 / 8 are not multiples of 16


def test_ffconvlstm_ctor_and_loud_failures():
    names = list(inspect.signature(FFConvLSTM.__init__).parameters)[1:16]
    assert names == ["in_dim", "ff_hidden_dim", "conv_hidden_dim", "lstm_hidden_dim", "out_dim", "dropout", "num_lstm_layers",
                     "bidirectional", "init_type", "use_mdn", "dim_wise", "num_gaussians", "in_ph_start_idx", "in_ph_end_idx",
                     "embed_dim"]
    assert list(inspect.signature(MultiSpeakerFFConvLSTM.__init__).parameters)[1:3] == ["in_dim", "speaker_embedding"]
    from ensemble_svs_with_interactions_b200.model import MDNLayer
    with pytest.raises(NotImplementedError):
        MDNLayer(8, 4, 3, dim_wise=False)
    gm = Golden("ffconvlstm_mdn")
    mm = FFConvLSTM(**gm.cfg)
    _same_layout(mm, gm.sd)                           # fc.log_pi / fc.log_sigma / fc.mu, in the reference's order
    assert mm.prediction_type() == PredictionType.PROBABILISTIC
    m = FFConvLSTM(87, ff_hidden_dim=64, conv_hidden_dim=32, lstm_hidden_dim=16, out_dim=32, in_ph_start_idx=3, in_ph_end_idx=50,
                   embed_dim=32, init_type="kaiming_normal")
    assert m.resolved_precision() == "bf16"
    assert torch.count_nonzero(m.fc.bias) == 0        # init_weights zeroes Linear / Conv biases (util.py:60-61)
    with pytest.raises(RuntimeError, match="no backward kernels"):      # training mode + trainable parameters + grad mode
        m(torch.zeros(1, 8, 87))
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):    # the training-mode forward itself is a CUDA path too
        m(torch.zeros(1, 8, 87))
    md = FFConvLSTM(87, ff_hidden_dim=64, conv_hidden_dim=32, lstm_hidden_dim=16, out_dim=32, dropout=0.1)
    with torch.no_grad(), pytest.raises(RuntimeError, match="dropout = 0 only"):
        md(torch.zeros(1, 8, 87))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.eval()(torch.zeros(1, 8, 87))


def test_postprocess_filter_design_matches_scipy():
    """Host side of row f3: the Butterworth design and lfilter_zi the device filter is fed with equal scipy's (dsp.py:25)."""
    from scipy import signal
    from ensemble_svs_with_interactions_b200.postprocess import butter_lowpass, lfilter_zi, lowpass_filter
    for N, Wn in ((5, 0.5), (5, 0.2), (3, 0.05), (8, 0.7)):
        b, a = butter_lowpass(N, Wn)
        bs, as_ = signal.butter(N, [Wn], "lowpass")
        assert np.abs(np.array(b) - bs).max() <= 1e-12 and np.abs(np.array(a) - as_).max() <= 1e-10
        assert np.abs(lfilter_zi(b, a) - signal.lfilter_zi(bs, as_)).max() <= 1e-9
    with pytest.raises(ValueError):
        butter_lowpass(5, 1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        lowpass_filter(torch.zeros(1, 40, 3), 200, cutoff=50)


def test_pipeline_pair_items_matches_reference_double_loop():
    """Row f3: same pairs, same order as the scan of nnsvs/bin/synthesis_multitrack.py:113-118."""
    from ensemble_svs_with_interactions_b200.pipeline import pair_items
    ids = ["alto_song1_seg0", "bass_song1_seg0", "alto_song1_seg1", "sop_song1_seg0", "bass_song2_seg0", "ritsu_song1_seg1", "solo"]
    brute = [(a, b) for a in ids for b in ids if a.split("_")[1:] == b.split("_")[1:]]
    assert pair_items(ids) == brute
    assert ("alto_song1_seg0", "alto_song1_seg0") in brute and ("solo", "solo") in brute and len(brute) == 3 * 3 + 2 * 2 + 1 + 1
    assert pair_items([]) == []


def test_wgrad_splits_choice():
    """Track groups of a wgrad launch: the largest divisor of B that keeps the grid within about one wave."""
    from ensemble_svs_with_interactions_b200 import ops
    assert ops.wgrad_splits(6, 36) == 3          # 108 CTAs; 6 groups would be 216
    assert ops.wgrad_splits(6, 12) == 6          # 72 CTAs
    assert ops.wgrad_splits(6, 200) == 1         # already more than a wave
    assert ops.wgrad_splits(1, 4) == 1
    assert ops.wgrad_splits(48, 36) == 4         # divisors of 48: 1 2 3 4 6 ...; 4 * 36 = 144 <= 160 < 6 * 36


def test_training_indicator_rows():
    """Rows appended to the wgrad operand: column sums over all / the first d / the last d frames of each track."""
    from ensemble_svs_with_interactions_b200.diffsinger import training
    ind = training._indicator_rows(2, 10, 16, 3, torch.device("cpu")).float()
    assert ind.shape == (2, 16, 16)
    assert ind[0, 0].tolist() == [1.0] * 10 + [0.0] * 6 and ind[0, 1, :4].tolist() == [1, 1, 1, 0]
    assert ind[0, 2].nonzero().flatten().tolist() == [7, 8, 9] and float(ind[0, 3:].abs().sum()) == 0   # track 0 owns rows 0..2
    assert ind[1, 3].sum() == 10 and float(ind[1, :3].abs().sum()) == 0                                 # track 1 owns rows 3..5
    ones = training._indicator_rows(2, 10, 16, None, torch.device("cpu")).float()
    assert ones.shape == (2, 16, 16) and ones[:, 0, :10].min() == 1 and float(ones[:, 1:].abs().sum()) == 0


def test_reference_root_discovery(tmp_path, monkeypatch):
    """oracle/ref_shim.py looks for the reference under $SVSK_REFERENCE_ROOT, baseline/_ref (the pip-installed copy that
    travels to the GPU box), then /root/reference."""
    import importlib
    from oracle import ref_shim
    fake = tmp_path / "ref"
    (fake / "nnsvs" / "diffsinger").mkdir(parents=True)
    monkeypatch.setenv("SVSK_REFERENCE_ROOT", str(fake))
    assert importlib.reload(ref_shim).REFERENCE_ROOT == str(fake)
    monkeypatch.delenv("SVSK_REFERENCE_ROOT")
    root = importlib.reload(ref_shim).REFERENCE_ROOT
    assert root.endswith("baseline/_ref") or root == "/root/reference"


@pytest.mark.parametrize("scales,Tf", [([5, 4, 3, 2], 23), ([4, 4, 4], 40), ([8, 8], 17), ([5, 4, 3, 2], 2)])
def test_usfgan_frame_window_covers_upsampler_reach(scales, Tf):
    """Frame-rate aux projection (csrc/usfgan_fr.cuh): for every 128-sample tile the 16-frame window that starts at
    svsk_usfgan_frame_base holds every frame the upsampler lets the tile hear, and 16 impulse channels (frame mod 16)
    tell those frames apart — checked against the oracle's upsampler on the CPU (the C function runs on the host)."""
    import math
    from ensemble_svs_with_interactions_b200 import _lib as L, ops
    from oracle import svs_oracle as O
    hop = int(np.prod(scales))
    reach, rate = 0, 1
    for s_ in scales:
        rate *= s_
        reach += s_ * (hop // rate)
    assert ops.usfgan_frame_window_ok(hop, reach)
    g = torch.Generator().manual_seed(Tf)
    sd = {"conv_in.weight": torch.zeros(1, 1, 1)}
    for n, s_ in enumerate(scales):
        sd[f"upsample.up_layers.{2 * n + 1}.weight"] = torch.rand(1, 1, 1, 2 * s_ + 1, generator=g) + 0.1

    def up(c):     # the oracle's stages without conv_in
        sd["conv_in.weight"] = torch.eye(c.shape[1]).unsqueeze(-1)
        return O.usfgan_upsample(sd, "", c, scales)
    c = torch.randn(1, 5, Tf, generator=g, dtype=torch.float64)
    full = up(c.float())[0].double()                                        # [5, T]
    T = Tf * hop
    imp = up((torch.arange(Tf)[None, :] % 16 == torch.arange(16)[:, None]).float()[None])[0].double()   # [16, T]
    fbase = L.lib().svsk_usfgan_frame_base
    assert fbase(0, reach, hop) == (math.floor(-reach / hop) // 8) * 8
    cpad = torch.zeros(5, 64 + Tf + 64, dtype=torch.float64)
    cpad[:, 64:64 + Tf] = c[0]
    for t0 in range(0, T, 128):
        n = min(128, T - t0)
        fb = fbase(t0, reach, hop)
        assert fb % 8 == 0 and fb == (math.floor((t0 - reach) / hop) // 8) * 8
        assert (t0 + 127 + reach) // hop - fb <= 15
        U = imp[[(fb + k) % 16 for k in range(16)]][:, t0:t0 + n]           # [16, n]
        rebuilt = cpad[:, 64 + fb:64 + fb + 16] @ U                          # [5, n]
        assert float((rebuilt - full[:, t0:t0 + n]).abs().max()) < 1e-5 * max(1.0, float(full.abs().max()))
    assert not ops.usfgan_frame_window_ok(12, 3 * 4 + 4)                     # hop 12 (scales [4, 3]): sample-rate path


def test_hoisted_conditioner_projection_host_checks():
    """Host side of the precomputed conditioner projection (ops.diffnet_pcond_pack / diffnet_stack_bf16(pcond=...)): shapes
    are checked before anything touches the device, and CPU tensors raise (no CPU path)."""
    from ensemble_svs_with_interactions_b200 import ops
    with pytest.raises(ValueError, match="p must be"):
        ops.diffnet_pcond_pack(torch.zeros(3, 10, 256, dtype=torch.bfloat16), 2, 5, 2, 256)      # L * 2C/256 = 4 blocks expected
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.diffnet_pcond_pack(torch.zeros(4, 10, 256, dtype=torch.bfloat16), 2, 5, 2, 256)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.diffnet_cond_project(torch.zeros(2, 5, 64, dtype=torch.bfloat16), torch.zeros(4, 256, 64, dtype=torch.bfloat16))
