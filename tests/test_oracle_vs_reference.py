"""Re-checks the CPU oracle against the LIVE reference modules on fresh seeded inputs (other shapes / seeds than the
committed fixtures).  Only runs where the reference tree is mounted (the build container); on the GPU box, where
``/root/reference`` does not exist, the whole module is skipped.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import svs_oracle as O
from oracle.ref_shim import load_reference, reference_available
from tests.golden_util import max_abs

pytestmark = pytest.mark.skipif(not reference_available() or torch.cuda.is_available(),
                                reason="needs the reference tree and a CUDA-free process (index.py calls .cuda())")


@pytest.fixture(scope="module")
def ns():
    torch.set_num_threads(1)
    return load_reference()


def close(a, b, tol=2e-5):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert max_abs(a, b) <= tol * max(1.0, b.abs().max().item()), max_abs(a, b)


def test_diffnet_live(ns):
    torch.manual_seed(101)
    cfg = dict(in_dim=12, encoder_hidden_dim=20, residual_layers=5, residual_channels=16, dilation_cycle_length=4)
    m = ns.DiffNet(**cfg).eval()
    with torch.no_grad():
        m.output_projection.weight.normal_(0, 0.2)
        spec, cond, t = torch.randn(2, 1, 12, 29), torch.randn(2, 20, 29), torch.tensor([5, 77])
        ref = m(spec, t, cond)
    close(O.diffnet_forward(m.state_dict(), spec, t, cond, 5, 4), ref)


def test_ffconvlstm_live(ns):
    torch.manual_seed(102)
    cfg = dict(in_dim=33, in_ph_start_idx=2, in_ph_end_idx=12, embed_dim=16, ff_hidden_dim=24, conv_hidden_dim=12,
               lstm_hidden_dim=12, num_lstm_layers=2, out_dim=7)
    m = ns.FFConvLSTM(**cfg).eval()
    for k, v in m.state_dict().items():
        if k.endswith("running_var"):
            v.copy_(torch.rand(v.shape) + 0.5)
        elif k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape) * 0.3)
    x = torch.randn(3, 21, 33)
    x[..., 2:12] = torch.nn.functional.one_hot(torch.randint(0, 10, (3, 21)), 10).float()
    lengths = [21, 20, 6]
    with torch.no_grad():
        ref = m(x.clone(), lengths)
    got = O.ffconvlstm_forward(m.state_dict(), x, lengths, in_ph_start_idx=2, in_ph_end_idx=12, embed_dim=16)
    close(got, ref)


def test_postprocess_live(ns):
    g = np.random.RandomState(103)
    x = np.cumsum(g.randn(333, 4), axis=0)
    for cutoff in (50, 20, 5):
        ref = np.stack([ns.lowpass_filter(x[:, d], 200, cutoff=cutoff) for d in range(4)], 1)
        got = np.stack([O.lowpass_filter(x[:, d], 200, cutoff=cutoff) for d in range(4)], 1)
        assert np.abs(got - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
    gv = g.rand(4) + 0.5
    idx = np.sort(g.choice(333, 200, replace=False))
    assert np.abs(O.variance_scaling(gv, x, 1, idx) - ns.variance_scaling(gv, x, offset=1, note_frame_indices=idx)).max() <= 1e-12


def test_wavenet_live(ns):
    torch.manual_seed(104)
    m = ns.WaveNet(in_dim=9, out_dim=6, layers=3, stacks=1, residual_channels=8, gate_channels=16, skip_out_channels=8,
                   kernel_size=2).eval()
    c, x = torch.rand(2, 17, 9), torch.rand(2, 17, 6)
    with torch.no_grad():
        ref = m(c, x)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    close(O.wavenet_forward(sd, c, x, layers=3, stacks=1), ref)


def test_parallel_hn_usfgan_live(ns):
    """Another seed, another shape than tests/golden/usfgan_parallel_hn_small.npz (aux 10, scales [2, 2], hop 4)."""
    torch.manual_seed(105)
    g = torch.Generator().manual_seed(106)
    common = dict(residual_channels=8, gate_channels=16, skip_channels=8, aux_channels=10, aux_context_window=2,
                  use_weight_norm=True, upsample_params={"upsample_scales": [2, 2]})
    pe = {"conv_layers": 3, "kernel_size": 5, "dilation": 1, "padding_mode": "replicate", "residual_channels": 8}
    hp = {"blockA": 3, "cycleA": 3, "blockF": 0, "cycleF": 0, "cascade_mode": 0}
    np_ = {"blockA": 0, "cycleA": 0, "blockF": 2, "cycleF": 1, "cascade_mode": 0}
    fp = {"blockA": 0, "cycleA": 0, "blockF": 4, "cycleF": 2, "cascade_mode": 0}
    m = ns.ParallelHnUSFGANGenerator(harmonic_network_params=dict(hp), noise_network_params=dict(np_),
                                     filter_network_params=dict(fp), periodicity_estimator_params=dict(pe), **common).eval()
    with torch.no_grad():
        m.periodicity_estimator.layers[-2].weight_v.normal_(0, 0.2, generator=g)
    B, Fr, hop = 2, 30, 4
    c = torch.randn(B, 10, Fr + 4, generator=g)
    f0 = torch.empty(B, Fr).uniform_(20.0, 60.0, generator=g)
    f0[:, 5:9] = 0.0
    d = torch.tensor(np.stack([ns.dilated_factor(f.numpy().astype(np.float64).copy(), 240, 4) for f in f0]),
                     dtype=torch.float32).repeat_interleave(hop, dim=-1)[:, None]
    x = torch.randn(B, 2, Fr * hop, generator=g) * 0.3
    with torch.no_grad():
        ref = m(x, c, d)
    pe_o = dict(pe); pe_o.pop("residual_channels")
    outs = O.parallel_hn_usfgan_forward(m.state_dict(), x, c, d, harmonic=hp, noise=np_, filt=fp, upsample_scales=[2, 2], pe=pe_o)
    for o, r in zip(outs, ref):
        close(o, r, 5e-5)
