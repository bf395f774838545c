"""Parity at the BASELINE configs in full, against the CPU oracle (oracle/svs_oracle.py), fp32 and bf16 stated separately.

The other GPU parity tests meet the oracle at sizes it finishes in a second or two; these are the headline shapes
themselves (VERDICT r1 "weak" #1: parity was only transitive there):

  config 2  DiffNet(80, 256, L20, C256), K = 100, 6 tracks x 2000 frames — injected x_T / z, per-step eps_hat at
            t in {99, 50, 0} and the final mel                           (nnsvs/diffsinger/diffusion.py:302-336)
  config 3  the recipe's ParallelHnUSFGANGenerator (20 adaptive + 5 + 30 fixed blocks, aux 80), one track x 5 s, with
            the oracle's in_signal, and through USFGANWrapper.inference   (nnsvs/usfgan/models/generator.py:472-522,
                                                                          nnsvs/usfgan/__init__.py:13-65)
  config 4  one (song, track) item through EnsembleSynthesizer (mgc + bap diffusion, K = 100, then the vocoder at aux 65)
            against the same chain of oracle functions
  a16 / a18 USFGANWrapper.inference against the reference's own output; WaveNet incremental == parallel logits

Tolerances are the assertions below; the measured errors are printed (pytest -s) and recorded in DESIGN.md §5.
Each oracle run costs tens of seconds of host CPU; the module-scoped fixtures run each of them once.
"""
import math
import time
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from oracle import svs_oracle as O
from tests.golden_util import Golden, max_abs, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def errs(a, b):
    """(relative L2, max-abs / max|ref|) of a against the reference b, both moved to the CPU."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.isfinite(a).all()
    return rel_l2(a, b), max_abs(a, b) / max(b.abs().max().item(), 1e-30)


def check(tag, a, b, l2, mx):
    r, m = errs(a, b)
    print(f"[fullsize] {tag}: rel_l2={r:.3e} max_abs/|ref|max={m:.3e}  (tolerance {l2:.0e} / {mx:.0e})")
    assert r <= l2 and m <= mx, (tag, r, m)
    return r, m


def random_diffnet(C, H, M, L, seed, cycle=4):
    from ensemble_svs_with_interactions_b200.diffsinger import DiffNet
    torch.manual_seed(seed)
    m = DiffNet(in_dim=M, encoder_hidden_dim=H, residual_layers=L, residual_channels=C, dilation_cycle_length=cycle)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        # zero-initialised in the reference (denoiser.py:99): parity would be vacuous
        m.output_projection.weight.copy_(torch.randn(m.output_projection.weight.shape, generator=g) * 0.05)
        for p in m.parameters():
            if p.dim() == 1:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    return m.eval()


def cpu_sd(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


# ------------------------------------------------------------------------------------------------ config 2
STEPS = (99, 50, 0)


@pytest.fixture(scope="module")
def config2():
    from ensemble_svs_with_interactions_b200.diffsinger import GaussianDiffusion
    B, T, M, H, C, L, K = 6, 2000, 80, 256, 256, 20, 100
    den = random_diffnet(C, H, M, L, seed=1234)
    m = GaussianDiffusion(H, M, den, K_step=K).eval()
    g = torch.Generator().manual_seed(1234)
    cond = torch.randn(B, T, H, generator=g)
    x_T = torch.randn(B, 1, M, T, generator=g)
    z = torch.randn(K, B, 1, M, T, generator=g)
    t0 = time.time()
    ref, traj = O.diffusion_inference(cpu_sd(m), cond, x_T, z, K_step=K, residual_layers=L, dilation_cycle_length=4,
                                      keep_steps=STEPS)
    print(f"[fullsize] config 2 oracle: {time.time() - t0:.1f} s on {torch.get_num_threads()} host threads")
    return NS(m=m.to(DEV), cond=cond.to(DEV), x_T=x_T.to(DEV), z=z.to(DEV), ref=ref, traj=traj)


def test_config2_bf16_sampling_vs_oracle(config2):
    """bf16 tensor-core path (what bench.py times): 100 chained steps of bf16 operands / tanh.approx against the fp32
    oracle.  The injected noise keeps both trajectories on the same path; the error that accumulates is stated here."""
    c = config2
    c.m.denoise_fn.precision = "auto"
    assert c.m.denoise_fn.resolved_precision() == "bf16"
    trace = {t: None for t in STEPS}
    y = c.m.inference(c.cond, x_T=c.x_T, z=c.z, trace=trace)
    check("config 2 bf16  eps_hat t=99", trace[99], c.traj[99][0], 2e-2, 3e-2)
    check("config 2 bf16  eps_hat t=50", trace[50], c.traj[50][0], 2e-2, 3e-2)
    check("config 2 bf16  eps_hat t=0 ", trace[0], c.traj[0][0], 2e-2, 3e-2)
    check("config 2 bf16  final mel   ", y, c.ref, 1.5e-2, 1.5e-1)
    # the CUDA-graph replay bench.py measures gives the same numbers as the traced eager launches, bit for bit
    assert torch.equal(c.m.inference(c.cond, x_T=c.x_T, z=c.z), y)


def test_config2_fp32_sampling_vs_oracle(config2):
    c = config2
    c.m.denoise_fn.precision = "fp32"
    try:
        trace = {t: None for t in STEPS}
        y = c.m.inference(c.cond, x_T=c.x_T, z=c.z, trace=trace)
    finally:
        c.m.denoise_fn.precision = "auto"
    check("config 2 fp32  eps_hat t=99", trace[99], c.traj[99][0], 1e-5, 2e-5)
    check("config 2 fp32  eps_hat t=50", trace[50], c.traj[50][0], 1e-5, 2e-5)
    check("config 2 fp32  eps_hat t=0 ", trace[0], c.traj[0][0], 1e-5, 2e-5)
    check("config 2 fp32  final mel   ", y, c.ref, 1e-5, 2e-5)


# ------------------------------------------------------------------------------------------------ config 3
PE = {"conv_layers": 3, "kernel_size": 5, "dilation": 1, "padding_mode": "replicate"}
HARMONIC = {"blockA": 20, "cycleA": 4, "blockF": 0, "cycleF": 0, "cascade_mode": 0}
NOISE = {"blockA": 0, "cycleA": 0, "blockF": 5, "cycleF": 5, "cascade_mode": 0}
FILTER = {"blockA": 0, "cycleA": 0, "blockF": 30, "cycleF": 3, "cascade_mode": 0}


class Cfg(dict):
    __getattr__ = dict.__getitem__


def vocoder_config(noise_amp=0.003, signal_types=("sine", "noise")):
    return NS(data=NS(sample_rate=24000, hop_size=120, sine_amp=0.1, noise_amp=noise_amp, signal_types=list(signal_types),
                      sine_f0_type="contf0", df_f0_type="contf0", dense_factor=4),
              generator=Cfg(aux_context_window=2))


def synthetic_f0(frames, gen, frame_rate=200):
    """piecewise-constant notes U(110, 880) Hz of 0.25-1 s, 15 % rests (f0 = 0), 5.5 Hz vibrato of
    +-30 cents."""
    f0 = torch.zeros(frames)
    t = 0
    while t < frames:
        n = int(torch.randint(frame_rate // 4, frame_rate + 1, (1,), generator=gen))
        if float(torch.rand(1, generator=gen)) >= 0.15:
            f0[t:t + n] = float(torch.empty(1).uniform_(110.0, 880.0, generator=gen))
        t += n
    vib = 2.0 ** (30.0 / 1200.0 * torch.sin(2 * math.pi * 5.5 * torch.arange(frames) / frame_rate))
    return (f0 * vib).float().numpy()[:, None]


def recipe_generator(aux_channels, seed):
    """The recipe's vocoder (conf/train_usfgan/generator/parallel_hn_usfgan.yaml: 20A cycle 4 / 5F cycle 5 / 30F cycle 3,
    64/128/64) with weight norm removed, biases and the periodicity estimator's last conv re-randomised (its 1e-4 init
    gives a = 0.5 everywhere)."""
    from ensemble_svs_with_interactions_b200.usfgan.models import ParallelHnUSFGANGenerator
    torch.manual_seed(seed)
    m = ParallelHnUSFGANGenerator(harmonic_network_params=dict(HARMONIC), noise_network_params=dict(NOISE),
                                  filter_network_params=dict(FILTER), periodicity_estimator_params=dict(PE),
                                  aux_channels=aux_channels).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        last = m.periodicity_estimator.layers[-2]
        last.weight_v.copy_(torch.randn(last.weight_v.shape, generator=g) * 0.1)
    m.remove_weight_norm()
    return m


def oracle_vocoder(sd, f0, aux, n_sine, n_in):
    x, c, d = O.usfgan_wrapper_inputs(f0, aux, sample_rate=24000, hop_size=120, dense_factor=4, aux_context_window=2,
                                      sine_amp=0.1, noise_amp=0.003, noise_sine=n_sine, noise_in=n_in)
    y = O.parallel_hn_usfgan_forward(sd, x, c, d, harmonic=HARMONIC, noise=NOISE, filt=FILTER, pe=PE)[0]
    return y, (x, c, d)


@pytest.fixture(scope="module")
def config3():
    frames = 1000                                             # 5 s at 24 kHz / hop 120 = 120 000 samples
    gen = recipe_generator(80, seed=4321)
    g = torch.Generator().manual_seed(1234)
    f0 = synthetic_f0(frames, g)
    aux = torch.randn(frames, 80, generator=g)
    n_sine = torch.randn(1, 1, frames * 120, generator=g)
    n_in = torch.randn(1, 1, frames * 120, generator=g)
    t0 = time.time()
    ref, (x, c, d) = oracle_vocoder(cpu_sd(gen), f0, aux, n_sine, n_in)
    print(f"[fullsize] config 3 oracle (55 blocks, {frames * 120} samples): {time.time() - t0:.1f} s; "
          f"|wav|max={ref.abs().max():.3f} rms={ref.pow(2).mean().sqrt():.3f}")
    return NS(gen=gen.to(DEV), f0=f0, aux=aux, n_sine=n_sine, n_in=n_in, ref=ref, x=x, c=c, d=d)


def test_config3_generator_bf16_and_fp32_vs_oracle(config3):
    """All 55 blocks with the oracle's own in_signal / aux / dilation factors: the stacks in isolation."""
    c = config3
    x, aux, d = c.x.to(DEV), c.c.to(DEV), c.d.to(DEV)
    c.gen.precision = "auto"
    assert c.gen._ntc_fast_path_ok()
    y = c.gen(x, aux, d, wave_only=True)[0]                   # the NTC bf16 path USFGANWrapper uses
    check("config 3 bf16  waveform (55 blocks, oracle in_signal)", y, c.ref, 1.5e-2, 3e-2)
    c.gen.precision = "fp32"
    try:
        y32 = c.gen(x, aux, d)[0]
    finally:
        c.gen.precision = "auto"
    check("config 3 fp32  waveform (55 blocks, oracle in_signal)", y32, c.ref, 5e-6, 1e-5)


def test_config3_wrapper_end_to_end_vs_oracle(config3):
    """The call the reference's callers make (gen.py:1694): numpy f0 + aux tensor in, waveform out — device-side source
    signal (fp32 cumsum phase over 120 000 samples) and dilation factors included."""
    from ensemble_svs_with_interactions_b200.usfgan import USFGANWrapper
    c = config3
    noise = {"sine": c.n_sine, "noise": c.n_in}
    w = USFGANWrapper(vocoder_config(), c.gen)
    c.gen.precision = "auto"
    y = w.inference(c.f0.copy(), c.aux.to(DEV), noise=noise)
    check("config 3 bf16  USFGANWrapper.inference", y, c.ref, 1.5e-2, 3e-2)
    yb = w.inference_batch(c.f0[None].copy(), c.aux[None].to(DEV), noise=noise)
    check("config 3 bf16  USFGANWrapper.inference_batch", yb, c.ref, 1.5e-2, 3e-2)
    c.gen.precision = "fp32"
    try:
        y32 = w.inference(c.f0.copy(), c.aux.to(DEV), noise=noise)
    finally:
        c.gen.precision = "auto"
    check("config 3 fp32  USFGANWrapper.inference", y32, c.ref, 5e-6, 1e-5)


# ------------------------------------------------------------------------------------------------ config 4
@pytest.fixture(scope="module")
def config4():
    """One (song, track) work item of BASELINE configs[3]: mgc DiffNet (M 60, H 256, L 20, C 256), bap DiffNet (M 5,
    C = H = 128, L 10; conf/train_acoustic/model/multitrack_acoustic_nnsvs_world_multi_ar_f0_diff_mgcbap.yaml:166-172),
    K = 100 each, then the recipe vocoder at aux 65 — 2000 frames = 10 s of audio."""
    from ensemble_svs_with_interactions_b200.diffsinger import GaussianDiffusion
    T, K = 2000, 100
    mgc = GaussianDiffusion(256, 60, random_diffnet(256, 256, 60, 20, seed=60), K_step=K).eval()
    bap = GaussianDiffusion(128, 5, random_diffnet(128, 128, 5, 10, seed=5), K_step=K).eval()
    gen = recipe_generator(65, seed=65)
    g = torch.Generator().manual_seed(4)
    cm, cb = torch.randn(T, 256, generator=g), torch.randn(T, 128, generator=g)
    f0 = synthetic_f0(T, g)
    nz = {"mgc": (torch.randn(1, 1, 60, T, generator=g), torch.randn(K, 1, 1, 60, T, generator=g)),
          "bap": (torch.randn(1, 1, 5, T, generator=g), torch.randn(K, 1, 1, 5, T, generator=g)),
          "vocoder": {"sine": torch.randn(1, 1, T * 120, generator=g), "noise": torch.randn(1, 1, T * 120, generator=g)}}
    t0 = time.time()
    m_ref = O.diffusion_inference(cpu_sd(mgc), cm[None], *nz["mgc"], K_step=K, residual_layers=20, dilation_cycle_length=4)
    b_ref = O.diffusion_inference(cpu_sd(bap), cb[None], *nz["bap"], K_step=K, residual_layers=10, dilation_cycle_length=4)
    aux_ref = torch.cat([m_ref, b_ref], dim=-1)[0]
    wav_ref, _ = oracle_vocoder(cpu_sd(gen), f0, aux_ref, nz["vocoder"]["sine"], nz["vocoder"]["noise"])
    print(f"[fullsize] config 4 oracle chain (1 item, {T} frames): {time.time() - t0:.1f} s")
    return NS(mgc=mgc.to(DEV), bap=bap.to(DEV), gen=gen.to(DEV), cm=cm, cb=cb, f0=f0, nz=nz, m_ref=m_ref, b_ref=b_ref,
              wav_ref=wav_ref[0, 0])


def _run_config4(c):
    from ensemble_svs_with_interactions_b200.pipeline import EnsembleSynthesizer
    from ensemble_svs_with_interactions_b200.usfgan import USFGANWrapper
    seen = {}

    def aux_fn(m, b, f0):
        seen["m"], seen["b"] = m.clone(), b.clone()
        return torch.cat([m, b], dim=-1)
    synth = EnsembleSynthesizer(c.mgc, c.bap, USFGANWrapper(vocoder_config(), c.gen), aux_fn=aux_fn)
    nz = {k: (tuple(t.to(DEV) for t in v) if isinstance(v, tuple) else v) for k, v in c.nz.items()}
    (wav,) = synth.synthesize([c.cm], [c.cb], [torch.from_numpy(c.f0)], noise=nz)
    return wav, seen["m"], seen["b"]


def test_config4_pipeline_item_bf16_vs_oracle(config4):
    c = config4
    wav, m, b = _run_config4(c)
    check("config 4 bf16  mgc stream (K=100)", m, c.m_ref, 1.5e-2, 1.5e-1)
    check("config 4 bf16  bap stream (K=100)", b, c.b_ref, 1.5e-2, 1e-1)
    check("config 4 bf16  waveform, whole chain", wav, c.wav_ref, 3e-2, 8e-2)


def test_config4_pipeline_item_fp32_vs_oracle(config4):
    c = config4
    mods = (c.mgc.denoise_fn, c.bap.denoise_fn, c.gen)
    for m_ in mods:
        m_.precision = "fp32"
    try:
        wav, m, b = _run_config4(c)
    finally:
        for m_ in mods:
            m_.precision = "auto"
    check("config 4 fp32  mgc stream (K=100)", m, c.m_ref, 5e-6, 2e-5)
    check("config 4 fp32  bap stream (K=100)", b, c.b_ref, 5e-6, 2e-5)
    check("config 4 fp32  waveform, whole chain", wav, c.wav_ref, 1e-5, 2e-5)


# ------------------------------------------------------------------------------------------------ a16: reference's own wrapper output
def _wrapper_from_golden(g):
    from ensemble_svs_with_interactions_b200.usfgan import USFGANWrapper
    from ensemble_svs_with_interactions_b200.usfgan.models import ParallelHnUSFGANGenerator
    cfg = g.cfg
    gen = ParallelHnUSFGANGenerator(harmonic_network_params=dict(cfg["harmonic"]), noise_network_params=dict(cfg["noise"]),
                                    filter_network_params=dict(cfg["filt"]), periodicity_estimator_params=dict(cfg["pe"]))
    gen.remove_weight_norm()
    gen.load_state_dict(g.sd, strict=True)
    return USFGANWrapper(vocoder_config(noise_amp=cfg["noise_amp"]), gen.to(DEV).eval()), gen


def test_usfgan_wrapper_inference_vs_reference_output():
    """Row a16 on the device: f0 (numpy) + aux -> waveform against what the UNMODIFIED reference wrapper returned for the
    same weights, inputs and Gaussian draws (tests/golden/usfgan_wrapper.npz, oracle/make_golden.py:golden_wrapper)."""
    g = Golden("usfgan_wrapper")
    w, gen = _wrapper_from_golden(g)
    noise = {"sine": g.inp["noise_sine"], "noise": g.inp["noise_in"]}
    f0, aux = g.inp["f0"].numpy(), g.inp["aux"].to(DEV)
    gen.precision = "fp32"
    check("a16 fp32  USFGANWrapper.inference vs reference", w.inference(f0.copy(), aux, noise=noise), g.out["wav"], 5e-6, 1e-5)
    gen.precision = "auto"
    assert gen._ntc_fast_path_ok()
    check("a16 bf16  USFGANWrapper.inference vs reference", w.inference(f0.copy(), aux, noise=noise), g.out["wav"], 2e-2, 3e-2)
    check("a16 bf16  inference_batch vs reference", w.inference_batch(f0[None].copy(), aux[None], noise=noise), g.out["wav"],
          2e-2, 3e-2)
    # the input construction itself, on the device, against the reference's tensors (tests/golden/usfgan_frontend.npz)
    from ensemble_svs_with_interactions_b200.usfgan.utils import SignalGenerator, dilated_factor
    fg = Golden("usfgan_frontend")
    fc = fg.cfg
    sg = SignalGenerator(sample_rate=fc["sample_rate"], hop_size=fc["hop_size"], sine_amp=fc["sine_amp"],
                         noise_amp=fc["noise_amp"], signal_types=["sine", "noise"])
    sg.injected_noise = {"sine": fg.inp["noise_sine"], "noise": fg.inp["noise_in"]}
    sig = sg(torch.FloatTensor(fg.inp["f0"].numpy()).unsqueeze(0).transpose(2, 1).to(DEV))
    # svsk_usfgan_source reproduces the reference's phase (fp64 prefix sums rounded to fp32) exactly; what is left is the
    # last-bit difference between the device's sinf and the host's sin, times the sine amplitude 0.1
    e = max_abs(sig.cpu(), fg.out["in_signal"])
    print(f"[fullsize] a16 source signal on the device vs reference tensors: max_abs={e:.3e}")
    assert sig.is_cuda and e <= 1e-6
    df = dilated_factor(np.squeeze(fg.inp["f0"].numpy().copy()), fc["sample_rate"], fc["dense_factor"]).repeat(fc["hop_size"])
    assert np.array_equal(df, fg.out["df"].numpy())


def test_source_signal_kernel_full_length_vs_oracle():
    """Row f2: the sine source and the dilation factors at BASELINE config 3's full length (6 tracks x 720 000 samples)
    against the oracle's sequential evaluation (features.py:145-164 through torch's CPU cumsum; features.py:56-75 in
    numpy float64).  SURVEY A.3.7 asks for a bound on the phase drift of a parallel scan at this length: the kernel scans
    in fp64 and rounds each prefix to fp32, exactly what the reference's CPU cumsum does, so there is no drift — the
    waveform differs only by the last bit of sin()."""
    from ensemble_svs_with_interactions_b200 import ops
    B, frames, hop, fs = 6, 6000, 120, 24000
    g = torch.Generator().manual_seed(99)
    f0 = np.stack([synthetic_f0(frames, g)[:, 0] for _ in range(B)]).astype(np.float32)          # [B, F]
    noise = torch.randn(B, 1, frames * hop, generator=g)
    ref = O.sine_source(torch.from_numpy(f0)[:, None], fs, hop, 0.1, 0.003, noise)
    d_ref = np.stack([O.dilated_factor(f0[b].copy(), fs, 4).repeat(hop) for b in range(B)]).astype(np.float32)
    f64 = torch.from_numpy(f0.astype(np.float64)).to(DEV)
    in_signal = torch.empty(B, 2, frames * hop, device=DEV)
    sine, d = ops.usfgan_source(f64, hop=hop, sample_rate=fs, dense_factor=4, sine_amp=0.1, noise_amp=0.003, noise=noise.to(DEV),
                                sine_out=in_signal)
    torch.cuda.synchronize()
    e = max_abs(in_signal[:, :1].cpu(), ref)
    print(f"[fullsize] source signal, 6 x 720000 samples: max_abs vs oracle = {e:.3e} (sine amplitude 0.1); "
          f"dilation factors bit-exact: {np.array_equal(d[:, 0].cpu().numpy(), d_ref)}")
    assert e <= 1e-6
    assert np.array_equal(d[:, 0].cpu().numpy(), d_ref)
    # float64 F0 from the caller (numpy default dtype): the factors follow the float64 values, the sine their fp32 rounding
    f0_64 = f0[:1].astype(np.float64) * (1 + 1e-9)
    _, d64 = ops.usfgan_source(torch.from_numpy(f0_64).to(DEV), hop=hop, sample_rate=fs, dense_factor=4, want_sine=False)
    assert np.array_equal(d64[0, 0].cpu().numpy(), O.dilated_factor(f0_64[0].copy(), fs, 4).repeat(hop).astype(np.float32))
    with pytest.raises(ValueError, match="Gaussian draws"):
        ops.usfgan_source(f64, hop=hop, sample_rate=fs, noise_amp=0.003)


# ------------------------------------------------------------------------------------------------ a18: incremental WaveNet
def test_wavenet_incremental_logits_equal_parallel_forward():
    """Row a18: teacher-forced, the incremental network (per-layer ring buffers, wavenet.py:117-139 / conv.py:21-53) must
    produce the logits of the parallel forward (libsvsk kernels) frame by frame — and both must equal the reference's."""
    from ensemble_svs_with_interactions_b200.wavenet import WaveNet
    g = Golden("wavenet_incremental")
    m = WaveNet(**g.cfg)
    m.load_state_dict(g.sd, strict=True)
    m = m.to(DEV).eval()
    c, x = g.inp["c"].to(DEV), g.inp["x"].to(DEV)
    par = m(c, x)
    m.clear_buffer()
    inc = torch.cat([m.incremental_logits(x[:, t:t + 1], c[:, t:t + 1]) for t in range(x.shape[1])], dim=1)
    m.clear_buffer()
    check("a18 incremental vs parallel logits (this repo)", inc, par, 5e-6, 5e-6)
    check("a18 parallel logits vs reference", par, g.out["parallel"], 5e-6, 5e-6)
    check("a18 incremental logits vs reference", inc, g.out["incremental"], 5e-6, 5e-6)
    # and the sampler built on it: one-hot frames, the right shape, buffers cleared afterwards
    torch.manual_seed(0)
    y = m.inference(c, num_time_steps=x.shape[1], tqdm=None)
    assert y.shape == x.shape and torch.equal(y.sum(-1), torch.ones_like(y.sum(-1)))
    assert m.first_conv.input_buffer is None
