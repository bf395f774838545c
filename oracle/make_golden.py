"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    CUDA_VISIBLE_DEVICES="" python oracle/make_golden.py

The reference's own tests hold no known-answer vectors for this path (shape checks
only: tests/test_wavenet.py:5-17, tests/test_diffusion.py:42-133), so the pins are
outputs of the reference modules themselves on seeded weights/inputs.  Each fixture
stores the reference ``state_dict`` (``sd/<key>``), the inputs (``in/<name>``), the
outputs (``out/<name>``) and the constructor config (``cfg`` json string), all fp32.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_shim import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _save(name, cfg, sd, ins, outs):
    blob = {"cfg": np.array(json.dumps(cfg))}
    for k, v in sd.items():
        blob["sd/" + k] = v.detach().cpu().numpy()
    for k, v in ins.items():
        blob["in/" + k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    for k, v in outs.items():
        blob["out/" + k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def _rerandomize(t, std, gen):
    with torch.no_grad():
        t.copy_(torch.randn(t.shape, generator=gen) * std)


def golden_diffnet(ns):
    torch.manual_seed(11)
    g = torch.Generator().manual_seed(12)
    cfg = dict(in_dim=16, encoder_hidden_dim=24, residual_layers=6, residual_channels=32, dilation_cycle_length=3)
    m = ns.DiffNet(**cfg).eval()
    _rerandomize(m.output_projection.weight, 0.2, g)  # zero-init in the reference (denoiser.py:99)
    for p in m.parameters():  # non-zero biases so bias paths are exercised
        if p.dim() == 1:
            _rerandomize(p, 0.1, g)
    B, T = 3, 37
    spec = torch.randn(B, 1, cfg["in_dim"], T, generator=g)
    t = torch.tensor([0, 3, 57])
    cond = torch.randn(B, cfg["encoder_hidden_dim"], T, generator=g)
    with torch.no_grad():
        y = m(spec, t, cond)
        # per-layer intermediates for kernel-level parity
        x = torch.relu(m.input_projection(spec[:, 0]))
        e = m.mlp(m.diffusion_embedding(t))
        x1, s1 = m.residual_layers[0](x, cond, e)
        x2, s2 = m.residual_layers[1](x1, cond, e)
    _save("diffnet_small", cfg, m.state_dict(), dict(spec=spec, t=t, cond=cond),
          dict(y=y, emb=e, x0=x, x1=x1, s1=s1, x2=x2, s2=s2))


def golden_diffusion(ns):
    torch.manual_seed(21)
    g = torch.Generator().manual_seed(22)
    dcfg = dict(in_dim=12, encoder_hidden_dim=20, residual_layers=4, residual_channels=16, dilation_cycle_length=2)
    K = 8
    den = ns.DiffNet(**dcfg)
    _rerandomize(den.output_projection.weight, 0.3, g)
    m = ns.GaussianDiffusion(in_dim=20, out_dim=12, denoise_fn=den, K_step=K).eval()
    B, T = 2, 24
    cond = torch.randn(B, T, 20, generator=g)
    x_T = torch.randn(B, 1, 12, T, generator=g)
    z = torch.randn(K, B, 1, 12, T, generator=g)
    # drive p_sample with injected noise: noise_fn is a p_sample kwarg
    x = x_T
    c = cond.transpose(1, 2)
    steps = []
    with torch.no_grad():
        for i in reversed(range(K)):
            t = torch.full((B,), i, dtype=torch.long)
            x = m.p_sample(x, t, c, noise_fn=lambda *s, device=None, _i=i: z[_i])
            steps.append(x)
        out = x[:, 0].transpose(1, 2) * m.norm_scale
    # training forward with injected t / noise: q_sample + denoise_fn as forward does (diffusion.py:288-299)
    y = torch.randn(B, T, 12, generator=g)
    tt = torch.tensor([1, 6])
    noise = torch.randn(B, 1, 12, T, generator=g)
    with torch.no_grad():
        xs = (y / m.norm_scale).transpose(1, 2)[:, None]
        x_noisy = m.q_sample(xs, tt, noise)
        eps = m.denoise_fn(x_noisy, tt, c)
    # PLMS transfer function (reachable by setting the attribute after construction)
    from collections import deque
    m.pndm_speedup = 2
    m.noise_list = deque(maxlen=4)
    with torch.no_grad():
        xp = x_T
        plms = []
        for i in reversed(range(0, K, 2)):
            xp = m.p_sample_plms(xp, torch.full((B,), i, dtype=torch.long), 2, c)
            plms.append(xp)
    m.pndm_speedup = None
    cfg = dict(denoiser=dcfg, K_step=K, in_dim=20, out_dim=12)
    _save("diffusion_small", cfg, m.state_dict(),
          dict(cond=cond, x_T=x_T, z=z, y=y, t_train=tt, noise_train=noise),
          dict(out=out, steps=torch.stack(steps), train_noise=noise.squeeze(1).transpose(1, 2),
               train_eps=eps.squeeze(1).transpose(1, 2), x_noisy=x_noisy, plms=torch.stack(plms)))


def golden_wavenet(ns):
    torch.manual_seed(31)
    g = torch.Generator().manual_seed(32)
    cfg = dict(in_dim=20, out_dim=12, layers=6, stacks=2, residual_channels=16, gate_channels=32,
               skip_out_channels=16, kernel_size=3)
    m = ns.WaveNet(**cfg).eval()
    for n, p in m.named_parameters():
        if n.endswith("weight_g"):
            _rerandomize(p, 1.0, g)  # g != ||v|| so the fold is exercised
            with torch.no_grad():
                p.abs_().add_(0.5)
        if n.endswith("bias"):
            _rerandomize(p, 0.1, g)
    B, T = 2, 50
    c = torch.rand(B, T, 20, generator=g)
    x = torch.rand(B, T, 12, generator=g)
    with torch.no_grad():
        y = m(c, x)
    _save("wavenet_small", cfg, m.state_dict(), dict(c=c, x=x), dict(y=y))
    # the reference test's own shape (tests/test_wavenet.py:5-17) at BASELINE config 1 batch
    torch.manual_seed(33)
    cfg2 = dict(in_dim=300, out_dim=206, layers=2, stacks=1, residual_channels=64, gate_channels=128,
                skip_out_channels=64, kernel_size=3)
    m2 = ns.WaveNet(**cfg2).eval()
    c2 = torch.rand(2, 200, 300, generator=g)
    x2 = torch.rand(2, 200, 206, generator=g)
    with torch.no_grad():
        y2 = m2(c2, x2)
    _save("wavenet_test_shape", cfg2, m2.state_dict(), dict(c=c2, x=x2), dict(y=y2))


def _f0_track(frames, g, lo=110.0, hi=440.0):
    f0 = torch.empty(frames).uniform_(lo, hi, generator=g)
    f0 = f0.repeat_interleave(5)[:frames]
    f0[frames // 3: frames // 3 + 4] = 0.0  # an unvoiced run
    return f0


def golden_usfgan(ns):
    g = torch.Generator().manual_seed(42)
    common = dict(residual_channels=16, gate_channels=32, skip_channels=16, aux_channels=12,
                  aux_context_window=2, use_weight_norm=True, upsample_params={"upsample_scales": [3, 2]})
    # NB the generators do not forward residual_channels to PeriodicityEstimator (generator.py:453-455):
    # its width defaults to 64 unless given here.
    pe = {"conv_layers": 3, "kernel_size": 5, "dilation": 1, "padding_mode": "replicate", "residual_channels": 16}
    B, Fr, hop = 2, 24, 6
    T = Fr * hop
    fs = 240  # tiny synthetic rate so dilations stay well inside T
    c = torch.randn(B, 12, Fr + 4, generator=g)
    f0 = torch.stack([_f0_track(Fr, g, 20.0, 60.0) for _ in range(B)])
    d = torch.tensor(np.stack([ns.dilated_factor(f.numpy().astype(np.float64).copy(), fs, 4) for f in f0]),
                     dtype=torch.float32).repeat_interleave(hop, dim=-1)[:, None]
    x2 = torch.randn(B, 2, T, generator=g) * 0.3
    x1 = x2[:, :1].contiguous()

    def shake(m):
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                _rerandomize(p, 0.1, g)
            if n.endswith("weight_g"):
                with torch.no_grad():
                    p.mul_(torch.empty(p.shape).uniform_(0.7, 1.3, generator=g))
        if hasattr(m, "periodicity_estimator"):
            last = m.periodicity_estimator.layers[-2]
            _rerandomize(last.weight_v, 0.2, g)  # std=1e-4 init gives a==0.5 everywhere (residual_block.py:381)
            with torch.no_grad():
                last.weight_g.fill_(1.5)

    # Parallel hn-uSFGAN
    torch.manual_seed(43)
    hp = {"blockA": 4, "cycleA": 2, "blockF": 0, "cycleF": 0, "cascade_mode": 0}
    np_ = {"blockA": 0, "cycleA": 0, "blockF": 2, "cycleF": 2, "cascade_mode": 0}
    fp = {"blockA": 0, "cycleA": 0, "blockF": 6, "cycleF": 2, "cascade_mode": 0}
    m = ns.ParallelHnUSFGANGenerator(harmonic_network_params=dict(hp), noise_network_params=dict(np_),
                                     filter_network_params=dict(fp), periodicity_estimator_params=dict(pe),
                                     **common).eval()
    shake(m)
    with torch.no_grad():
        y, s, h, n, a = m(x2, c, d)
    cfg = dict(harmonic=hp, noise=np_, filt=fp, pe=pe, common=common)
    _save("usfgan_parallel_hn_small", cfg, m.state_dict(), dict(x=x2, c=c, d=d), dict(y=y, s=s, h=h, n=n, a=a))
    # same weights after remove_weight_norm (the state the vocoder runs in, nnsvs/util.py:414)
    m.remove_weight_norm()
    with torch.no_grad():
        y_nw = m(x2, c, d)[0]
    _save("usfgan_parallel_hn_small_nowm", cfg, m.state_dict(), dict(x=x2, c=c, d=d), dict(y=y_nw))

    # Cascade hn-uSFGAN
    torch.manual_seed(44)
    mc = ns.CascadeHnUSFGANGenerator(harmonic_network_params=dict(hp), noise_network_params=dict(np_),
                                     filter_network_params=dict(fp), periodicity_estimator_params=dict(pe),
                                     **common).eval()
    shake(mc)
    with torch.no_grad():
        y, s, h, n, a = mc(x2, c, d)
    _save("usfgan_cascade_hn_small", cfg, mc.state_dict(), dict(x=x2, c=c, d=d), dict(y=y, s=s, h=h, n=n, a=a))

    # plain uSFGAN (source A -> filter F), cascade_mode 1 in the source net to cover F->A ordering
    torch.manual_seed(45)
    sp = {"blockA": 4, "cycleA": 2, "blockF": 2, "cycleF": 1, "cascade_mode": 1}
    mu = ns.USFGANGenerator(source_network_params=dict(sp), filter_network_params=dict(fp), **common).eval()
    shake(mu)
    with torch.no_grad():
        y, s = mu(x1, c, d)
    _save("usfgan_plain_small", dict(source=sp, filt=fp, common=common), mu.state_dict(),
          dict(x=x1, c=c, d=d), dict(y=y, s=s))

    # single blocks at the real channel widths (64/128/64, aux 80) incl. a large reflect dilation
    torch.manual_seed(46)
    fb = ns.FixedBlock(64, 128, 64, 80, kernel_size=3, dilation=64).eval()
    ab = ns.AdaptiveBlock(64, 128, 64, 80).eval()
    for blk in (fb, ab):
        for n_, p in blk.named_parameters():
            if n_.endswith("bias"):
                _rerandomize(p, 0.1, g)
    Tb = 160
    xb = torch.randn(1, 64, Tb, generator=g)
    cb = torch.randn(1, 80, Tb, generator=g)
    db = torch.empty(1, 1, Tb).uniform_(0.8, 9.7, generator=g)
    db[0, 0, :7] = torch.tensor([0.5, 1.5, 2.5, 3.5, 4.5, 5.5, 6.5])  # half-to-even ties
    bi, ci = ns.index_initial(1, 64)
    with torch.no_grad():
        yf, sf = fb(xb, cb)
        xP, xF = ns.pd_indexing(xb, db, 4, bi, ci)
        ya, sa = ab(xb, xP, xF, cb)
    sd = {"fixed." + k: v for k, v in fb.state_dict().items()}
    sd.update({"adaptive." + k: v for k, v in ab.state_dict().items()})
    _save("usfgan_blocks_fullwidth", dict(dilation_fixed=64, dilation_adaptive=4), sd,
          dict(x=xb, c=cb, d=db), dict(y_fixed=yf, y_adaptive=ya, xP=xP, xF=xF))


def golden_frontend(ns):
    """USFGANWrapper.inference input construction (nnsvs/usfgan/__init__.py:13-65)."""
    g = torch.Generator().manual_seed(52)
    Fr, hop, fs = 30, 12, 2400
    f0 = _f0_track(Fr, g, 100.0, 400.0).numpy().astype(np.float32)[:, None]
    sg = ns.SignalGenerator(sample_rate=fs, hop_size=hop, sine_amp=0.1, noise_amp=0.003,
                            signal_types=["sine", "noise"])
    f0_t = torch.FloatTensor(f0).unsqueeze(0).transpose(2, 1)
    torch.manual_seed(53)
    sig = sg(f0_t)
    # the same draws, in the reference's order: sine-noise first, then the noise channel
    torch.manual_seed(53)
    n_sine = torch.randn(1, 1, Fr * hop)
    n_in = torch.randn(1, 1, Fr * hop)
    df = ns.dilated_factor(np.squeeze(f0.copy()), fs, 4).repeat(hop, axis=0)
    _save("usfgan_frontend", dict(sample_rate=fs, hop_size=hop, dense_factor=4, sine_amp=0.1, noise_amp=0.003),
          {}, dict(f0=f0, noise_sine=n_sine, noise_in=n_in), dict(in_signal=sig, df=df.astype(np.float64)))


def golden_wrapper(ns):
    """The whole ``USFGANWrapper.inference`` call (nnsvs/usfgan/__init__.py:13-65) at the recipe's widths (64/128/64,
    aux 80, 24 kHz, hop 120, upsample 5*4*3*2) with short stacks: f0 + aux in, waveform out.  The two Gaussian draws of
    the source signal are recorded (the reference takes them from torch's global CPU generator)."""
    from types import SimpleNamespace as NS
    g = torch.Generator().manual_seed(71)
    torch.manual_seed(72)
    hp = {"blockA": 4, "cycleA": 2, "blockF": 0, "cycleF": 0, "cascade_mode": 0}
    np_ = {"blockA": 0, "cycleA": 0, "blockF": 2, "cycleF": 2, "cascade_mode": 0}
    fp = {"blockA": 0, "cycleA": 0, "blockF": 6, "cycleF": 3, "cascade_mode": 0}
    pe = {"conv_layers": 3, "kernel_size": 5, "dilation": 1, "padding_mode": "replicate"}
    m = ns.ParallelHnUSFGANGenerator(harmonic_network_params=dict(hp), noise_network_params=dict(np_),
                                     filter_network_params=dict(fp), periodicity_estimator_params=dict(pe)).eval()
    for n, p in m.named_parameters():
        if n.endswith("bias"):
            _rerandomize(p, 0.1, g)
    last = m.periodicity_estimator.layers[-2]
    _rerandomize(last.weight_v, 0.2, g)
    m.remove_weight_norm()

    class Cfg(dict):
        __getattr__ = dict.__getitem__
    fs, hop, Fr = 24000, 120, 40
    config = NS(data=NS(sample_rate=fs, hop_size=hop, sine_amp=0.1, noise_amp=0.003, signal_types=["sine", "noise"],
                        sine_f0_type="contf0", df_f0_type="contf0", dense_factor=4),
                generator=Cfg(aux_context_window=2))
    f0 = _f0_track(Fr, g, 110.0, 660.0).numpy().astype(np.float32)[:, None]
    aux = torch.randn(Fr, 80, generator=g)
    torch.manual_seed(73)
    with torch.no_grad():
        wav = ns.USFGANWrapper(config, m).inference(f0.copy(), aux)
    torch.manual_seed(73)  # the same draws, in the reference's order: the sine's additive noise, then the noise channel
    n_sine = torch.randn(1, 1, Fr * hop)
    n_in = torch.randn(1, 1, Fr * hop)
    _save("usfgan_wrapper", dict(harmonic=hp, noise=np_, filt=fp, pe=pe, sample_rate=fs, hop_size=hop, dense_factor=4,
                                 sine_amp=0.1, noise_amp=0.003, aux_context_window=2),
          m.state_dict(), dict(f0=f0, aux=aux, noise_sine=n_sine, noise_in=n_in), dict(wav=wav))


def golden_wavenet_incremental(ns):
    """Teacher-forced incremental evaluation (the loop body of WaveNet.inference, wavenet.py:117-139, fed a given
    one-hot sequence instead of its own samples) next to the parallel forward of the same network."""
    torch.manual_seed(81)
    g = torch.Generator().manual_seed(82)
    cfg = dict(in_dim=20, out_dim=12, layers=6, stacks=2, residual_channels=16, gate_channels=32, skip_out_channels=16,
               kernel_size=3)
    m = ns.WaveNet(**cfg).eval()
    for n, p in m.named_parameters():
        if n.endswith("bias"):
            _rerandomize(p, 0.1, g)
    B, T = 2, 30
    c = torch.randn(B, T, cfg["in_dim"], generator=g)
    x = torch.nn.functional.one_hot(torch.randint(0, cfg["out_dim"], (B, T), generator=g), cfg["out_dim"]).float()
    with torch.no_grad():
        par = m(c, x)
        m.clear_buffer()
        inc = []
        for t in range(T):
            h = m.first_conv.incremental_forward(x[:, t:t + 1])
            skips = 0
            for f in m.main_conv_layers:
                h, sk = f.incremental_forward(h, c[:, t:t + 1])
                skips = skips + sk
            h = skips
            for f in m.last_conv_layers:
                h = f.incremental_forward(h) if hasattr(f, "incremental_forward") else f(h)
            inc.append(h)
        m.clear_buffer()
    _save("wavenet_incremental", cfg, m.state_dict(), dict(c=c, x=x), dict(parallel=par, incremental=torch.cat(inc, dim=1)))


def golden_encoder(ns):
    """FFConvLSTM in eval mode (model.py:779-926): the recipe's embedding front with ragged lengths and all-zero phoneme
    blocks, and the shape of the reference's own test (tests/test_model.py:190-205)."""
    for name, cfg, B, T, lengths in (
        ("ffconvlstm_embed", dict(in_dim=60, in_ph_start_idx=3, in_ph_end_idx=20, embed_dim=24, ff_hidden_dim=32,
                                  conv_hidden_dim=24, lstm_hidden_dim=40, num_lstm_layers=2, out_dim=16), 3, 50, [50, 41, 17]),
        ("ffconvlstm_test_shape", dict(in_dim=300, ff_hidden_dim=8, conv_hidden_dim=8, lstm_hidden_dim=8, dropout=0.1,
                                       num_lstm_layers=2, bidirectional=True, out_dim=180, init_type="none"), 2, 33, [33, 29]),
        # the MDN head (tests/test_model.py builds the same model with use_mdn=True)
        ("ffconvlstm_mdn", dict(in_dim=40, ff_hidden_dim=32, conv_hidden_dim=16, lstm_hidden_dim=16, num_lstm_layers=2,
                                out_dim=11, use_mdn=True, dim_wise=True, num_gaussians=3), 2, 25, [25, 18]),
    ):
        torch.manual_seed(61)
        g = torch.Generator().manual_seed(62)
        m = ns.FFConvLSTM(**cfg).eval()
        for k, v in m.state_dict().items():
            if k.endswith("running_mean"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.2)
            elif k.endswith("running_var"):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
            elif ".bias" in k and v.dim() == 1 and "lstm" not in k:
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        for k in (2, 6, 10):  # BatchNorm scale away from 1
            m.conv[k].weight.data.copy_(torch.rand(m.conv[k].weight.shape, generator=g) + 0.5)
        x = torch.randn(B, T, cfg["in_dim"], generator=g)
        if cfg.get("embed_dim") is not None:
            s, V = cfg["in_ph_start_idx"], cfg["in_ph_end_idx"] - cfg["in_ph_start_idx"]
            ph = torch.randint(0, V, (B, T), generator=g)
            onehot = torch.nn.functional.one_hot(ph, V).float()
            onehot[:, ::7] = 0.0  # frames without a phoneme: argmax of zeros -> phoneme 0 (model.py:908)
            x[..., s:s + V] = onehot
        with torch.no_grad():
            y = m(x.clone(), lengths)
            extra = {}
            if cfg.get("use_mdn"):
                log_pi, log_sigma, mu = y
                mu_best, sigma_best = m.inference(x.clone(), lengths)     # (mu, sigma) of the most probable component
                extra = dict(log_pi=log_pi, log_sigma=log_sigma, mu=mu, mu_best=mu_best, sigma_best=sigma_best)
                y = mu
            ff = m.ff(x if cfg.get("embed_dim") is None else
                      m.emb(torch.argmax(x[..., s:s + V], -1)) + m.fc_in(torch.cat([x[..., :s], x[..., s + V:]], -1)))
            conv = m.conv(ff.transpose(1, 2)).transpose(1, 2)
        _save(name, cfg, m.state_dict(), dict(x=x, lengths=torch.tensor(lengths)), dict(y=y, ff=ff, conv=conv, **extra))


def golden_encoder_train(ns):
    """FFConvLSTM in TRAINING mode with dropout = 0 (the diffusion recipe's encoders, multitrack_..._diff_mgcbap.yaml:115,157):
    BatchNorm1d normalises with the batch statistics of the padded batch and updates its running buffers
    (model.py:839-852,915).  Two consecutive forwards, so that the second one starts from updated buffers."""
    cfg = dict(in_dim=60, in_ph_start_idx=3, in_ph_end_idx=20, embed_dim=24, ff_hidden_dim=32, conv_hidden_dim=24,
               lstm_hidden_dim=40, num_lstm_layers=2, out_dim=16, dropout=0.0)
    B, T, lengths = 3, 50, [50, 41, 17]
    torch.manual_seed(71)
    g = torch.Generator().manual_seed(72)
    m = ns.FFConvLSTM(**cfg).train()
    for k, v in m.state_dict().items():
        if k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.2)
        elif k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        elif ".bias" in k and v.dim() == 1 and "lstm" not in k:
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    for k in (2, 6, 10):
        m.conv[k].weight.data.copy_(torch.rand(m.conv[k].weight.shape, generator=g) + 0.5)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    s, V = cfg["in_ph_start_idx"], cfg["in_ph_end_idx"] - cfg["in_ph_start_idx"]
    xs = []
    for _ in range(2):
        x = torch.randn(B, T, cfg["in_dim"], generator=g)
        x[..., s:s + V] = torch.nn.functional.one_hot(torch.randint(0, V, (B, T), generator=g), V).float()
        xs.append(x)
    with torch.no_grad():
        y1 = m(xs[0].clone(), lengths)
        bn1 = {k: v.clone() for k, v in m.state_dict().items() if "running_" in k or "num_batches" in k}
        y2 = m(xs[1].clone(), lengths)
        bn2 = {k: v.clone() for k, v in m.state_dict().items() if "running_" in k or "num_batches" in k}
    out = dict(y1=y1, y2=y2)
    out.update({"bn1." + k: v.float() for k, v in bn1.items()})
    out.update({"bn2." + k: v.float() for k, v in bn2.items()})
    _save("ffconvlstm_train", cfg, sd0, dict(x1=xs[0], x2=xs[1], lengths=torch.tensor(lengths)), out)


def golden_postprocess(ns):
    """nnsvs.dsp.lowpass_filter and nnsvs.postfilters.variance_scaling as gen.postprocess_acoustic calls them
    (gen.py:1394-1418,1500-1513): 5 ms frames (modfs 200), cutoffs 50 (mgc / bap) and 20 (lf0)."""
    g = np.random.RandomState(71)
    T, D = 400, 7
    t = np.arange(T)[:, None] / 200.0
    x = np.sin(2 * np.pi * (3.0 + np.arange(D)[None]) * t) + 0.3 * g.randn(T, D)        # float64, like the scaler's output
    y50 = np.stack([ns.lowpass_filter(x[:, d], 200, cutoff=50) for d in range(D)], 1)
    y20 = np.stack([ns.lowpass_filter(x[:, d], 200, cutoff=20) for d in range(D)], 1)
    short = x[:18, 0].copy()
    y_short = ns.lowpass_filter(short, 200, cutoff=50)                                   # too short: returned as is
    y_19 = ns.lowpass_filter(x[:19, 0].copy(), 200, cutoff=50)
    gv = g.rand(D) + 0.5
    notes = np.sort(g.choice(T, 300, replace=False))
    vs_notes = ns.variance_scaling(gv, x, offset=2, note_frame_indices=notes)
    vs_all = ns.variance_scaling(gv, x, offset=2)
    vs_none = ns.variance_scaling(gv, x, offset=2, note_frame_indices=np.array([], dtype=np.int64))
    _save("postprocess", dict(fs=200, N=5), {}, dict(x=x, gv=gv, notes=notes, short=short),
          dict(y50=y50, y20=y20, y_short=y_short, y_19=y_19, vs_notes=vs_notes, vs_all=vs_all, vs_none=vs_none))


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = load_reference()
    assert not torch.cuda.is_available(), "run with CUDA_VISIBLE_DEVICES='' (index.py calls .cuda())"
    torch.set_num_threads(1)  # deterministic reduction order
    makers = [golden_diffnet, golden_diffusion, golden_wavenet, golden_usfgan, golden_frontend, golden_encoder,
              golden_encoder_train, golden_postprocess, golden_wrapper, golden_wavenet_incremental]
    only = set(sys.argv[1:])          # e.g. "golden_wrapper": regenerate that fixture only
    for mk in makers:
        if not only or mk.__name__ in only:
            mk(ns)


if __name__ == "__main__":
    main()
