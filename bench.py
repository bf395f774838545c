#!/usr/bin/env python
"""bench.py — one JSON line for the hot path's headline metric.

Workload (BASELINE.json configs[1]): DiffSinger DiffNet denoiser (20 residual layers, 256 ch, mel 80) run for a
100-step DDPM sampling over a 6-part ensemble batch x 2000 frames on one B200.  A "step" is ONE full 100-step
sampling pass over one batch; the metric is denoiser frame-steps per second (B*T*K_step per pass).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 (torchrun): every rank samples its own 6-track batch (tracks are independent work items, no data-path
collective) -> weak scaling; the time is the max over ranks.
--impl reference: the UNMODIFIED reference modules (baseline/_ref, loaded through oracle/ref_shim.py; the oracle port if
no reference tree travelled) on the box's host cores, fp32, all host threads, on a bounded sample of the same workload;
rank 0 only.

Besides the headline (config 2) the line carries, for every N: "vocoder" (config 3, one batch per rank), "pipeline"
(config 4: 64 songs split over the ranks by sharding.assign, STRONG scaling), "training" (config 5: DDP step, 6 x 1000
frames per rank); and at N = 1 also "wavenet" (config 1, GPU and reference CPU), "reference_gpu_eager" (the reference's
own modules in eager PyTorch on this GPU: the on-box bar) and "cpu_baseline".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(in_dim=80, encoder_hidden_dim=256, residual_layers=20, residual_channels=256, dilation_cycle_length=4)
K_STEP, B, T = 100, 6, 2000
MAC_PER_FRAME_STEP = 13_213_696          # reference formulation incl. conditioner projection
BLOCK_MAC_PER_FRAME = 655_360            # one fused block launch: 2C*(3C+H) + 2C*C per frame
METRIC = "denoiser frame-steps/sec (DiffNet 20x256, 100-step DDPM sampling)"
# identical in both arms (the driver compares it)
CONFIG = {"workload": f"DiffNet(80,256,L20,C256) 100-step DDPM sampling, {B} tracks x {T} frames per GPU",
          "sharding": "one 6-track batch per rank, no collective",
          "l2": "GPU arm: 256 MB flush between timed passes; the 384 MB noise tensor of a pass exceeds L2"}


STACK_NCU_SUMMARY = "r02q_stack_ncu_full_summary.json"        # dram bytes of diffnet_stack_kernel<hoisted projection> (ncu --set full)
STACK_NCU_SUMMARY_IN_GEMM = "r02k_stack_ncu_full_summary.json"
USFGAN_NCU_SUMMARY = "r02r_usfgan_block_fr_ncu_full_summary.json"


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def _ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` summary (or None)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", name)))[0]
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            k = [x for x in d if x.startswith(key + " [")][0]
            tot += float(d[k]) * unit[k[k.index("[") + 1:-1]]
        return tot
    except Exception:
        return None


def build_model(seed=1234):
    from ensemble_svs_with_interactions_b200.diffsinger import DiffNet, GaussianDiffusion
    torch.manual_seed(seed)
    den = DiffNet(**CFG)
    with torch.no_grad():
        den.output_projection.weight.normal_(0, 0.02)   # zero-init in the reference would make the work vacuous
    return GaussianDiffusion(CFG["encoder_hidden_dim"], CFG["in_dim"], den, K_step=K_STEP).eval()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index):
        self.rows, self.stop = [], threading.Event()
        self.cmd = ["nvidia-smi", "-i", str(index),
                    "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                    "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                    "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits"]
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(self.cmd, capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def load_reference_modules():
    """The unmodified reference classes (baseline/_ref or /root/reference) or None."""
    try:
        from oracle import ref_shim
        if ref_shim.reference_available():
            return ref_shim.load_reference()
    except Exception as e:      # a broken install must not take the bench down: fall back to the port, and say so
        print(f"bench.py: reference modules unavailable ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
    return None


def reference_model(ns, seed=1234):
    """nnsvs.diffsinger.GaussianDiffusion(DiffNet) at BASELINE config 2, built by the reference's own constructors."""
    torch.manual_seed(seed)
    den = ns.DiffNet(**CFG)
    with torch.no_grad():
        den.output_projection.weight.normal_(0, 0.02)
    return ns.GaussianDiffusion(in_dim=CFG["encoder_hidden_dim"], out_dim=CFG["in_dim"], denoise_fn=den, K_step=K_STEP).eval()


def reference_steps(device, sample_steps, repeats=1, threads=None):
    """`sample_steps` DDPM steps of the full B x T batch through the reference's stock code path (GaussianDiffusion.
    p_sample -> DiffNet.forward, diffusion.py:193-204) on `device`; the oracle port when no reference tree is present.
    Returns (frame-steps/s, kind, description, seconds per repeat, threads)."""
    threads = threads or os.cpu_count() or 1
    if device.type == "cpu":
        torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1234)
    cond = torch.randn(B, T, CFG["encoder_hidden_dim"], generator=g).transpose(1, 2).contiguous().to(device)
    x_T = torch.randn(B, 1, CFG["in_dim"], T, generator=g).to(device)
    ns = load_reference_modules()
    if ns is not None:
        m = reference_model(ns).to(device)
        kind = "reference"

        def run():
            x = x_T
            for i in reversed(range(K_STEP - sample_steps, K_STEP)):
                x = m.p_sample(x, torch.full((B,), i, device=device, dtype=torch.long), cond)
            return x
    else:
        from oracle import svs_oracle as O
        m = build_model()
        sd = {k: v.detach().to(device) for k, v in m.state_dict().items()}
        den = {k[len("denoise_fn."):]: v for k, v in sd.items() if k.startswith("denoise_fn.")}
        tab = {k: sd[k] for k in O.SCHEDULE_BUFFERS}
        kind = "port"

        def run():
            x = x_T
            for i in reversed(range(K_STEP - sample_steps, K_STEP)):
                t = torch.full((B,), i, dtype=torch.long, device=device)
                eps = O.diffnet_forward(den, x, t, cond, CFG["residual_layers"], CFG["dilation_cycle_length"])
                x = O.ddpm_update(tab, x, t, eps, torch.randn_like(x))
            return x

    def sync():
        if device.type == "cuda":
            torch.cuda.synchronize(device)

    with torch.no_grad():
        run()  # warm-up
        sync()
        t0 = time.perf_counter()
        for _ in range(repeats):
            run()
        sync()
        dt = (time.perf_counter() - t0) / repeats
    return (B * T * sample_steps / dt, kind,
            f"{sample_steps} of {K_STEP} DDPM steps of the full {B}x{T} batch (p_sample -> DiffNet.forward, fp32)", dt, threads)


def cpu_baseline(sample_steps=2, threads=None):
    v, kind, sample, dt, threads = reference_steps(torch.device("cpu"), sample_steps, threads=threads)
    return {"value": v, "unit": "frame-steps/s", "cores": threads, "kind": kind, "sample": sample, "seconds": dt}


def wavenet_bench(dev):
    """BASELINE configs[0]: nnsvs WaveNet forward at the shapes of the reference's tests/test_wavenet.py (in_dim 300,
    out_dim 206, layers 2), batch 2 x 200 frames, fp32 — this library on the GPU, the reference module on the host."""
    from ensemble_svs_with_interactions_b200.wavenet import WaveNet
    kw = dict(in_dim=300, out_dim=206, layers=2)
    Bw, Tw = 2, 200
    torch.manual_seed(1234)
    m = WaveNet(**kw).to(dev).eval()
    g = torch.Generator().manual_seed(1234)
    c, x = torch.rand(Bw, Tw, 300, generator=g), torch.rand(Bw, Tw, 206, generator=g)
    cd, xd = c.to(dev), x.to(dev)
    for _ in range(5):
        m(cd, xd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        m(cd, xd)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out = {"metric": "WaveNet forward frames/sec (in 300, out 206, layers 2, B 2 x T 200, fp32)", "value": Bw * Tw / (ms / 1e3),
           "unit": "frames/s", "ms_per_forward": ms, "dtype": "f32"}
    ns = load_reference_modules()
    if ns is not None:
        torch.manual_seed(1234)
        ref = ns.WaveNet(**kw).eval()
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            for _ in range(3):
                ref(c, x)
            t0 = time.perf_counter()
            for _ in range(20):
                ref(c, x)
            cpu_ms = (time.perf_counter() - t0) / 20 * 1e3
        out["cpu_reference"] = {"value": Bw * Tw / (cpu_ms / 1e3), "unit": "frames/s", "ms_per_forward": cpu_ms,
                                "cores": os.cpu_count(), "kind": "reference"}
    return out


def vocoder_bench(dev, peaks, tracks=6, seconds=30.0, reps=10):
    """Second half of the headline metric (BASELINE configs[2]): ParallelHn-uSFGAN, recipe config, 6 tracks x 30 s at
    24 kHz, residual stacks on the fused tcgen05 block kernel.  Returns a dict for the JSON line."""
    from ensemble_svs_with_interactions_b200 import ops
    from ensemble_svs_with_interactions_b200.usfgan.models import ParallelHnUSFGANGenerator
    fs, hop = 24000, 120
    frames = int(seconds * fs / hop)
    Tn = frames * hop
    pe = {"conv_layers": 3, "kernel_size": 5, "dilation": 1, "padding_mode": "replicate"}
    torch.manual_seed(1234)
    m = ParallelHnUSFGANGenerator(periodicity_estimator_params=pe).eval()
    with torch.no_grad():
        m.periodicity_estimator.layers[-2].weight_v.normal_(0, 0.05)
    m.remove_weight_norm()
    m = m.to(dev)
    g = torch.Generator().manual_seed(1)
    c = torch.randn(tracks, 80, frames + 4, generator=g).to(dev)
    f0 = torch.empty(tracks, 1, frames).uniform_(110, 880, generator=g)
    d = (fs / (f0 * 4)).repeat_interleave(hop, dim=-1).to(dev)
    x = (torch.randn(tracks, 2, Tn, generator=g) * 0.1).to(dev)
    for _ in range(3):
        m(x, c, d, wave_only=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index or 0) as clk:      # the HBM-heavy block kernel is the one a power cap shows on first
        e0.record()
        for _ in range(reps):
            m(x, c, d, wave_only=True)
        e1.record()
        e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # dominant kernel of the vocoder: the fused block with the frame-rate aux projection, as the pass runs it (the 30
    # fixed blocks of the filter network timed in isolation)
    frames_op = m._aux_frames(c, Tn, [m.filter_network])
    auxb = None if frames_op is not None else ops.nct_to_ntc(m.upsample_net(c), Cp=80)[0]
    hb = torch.randn(tracks, Tn, 64, device=dev).to(torch.bfloat16)
    m.filter_network.forward_ntc_bf16(hb, auxb, d, {}, frames=frames_op)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        m.filter_network.forward_ntc_bf16(hb, auxb, d, {}, frames=frames_op)
    e1.record()
    e1.synchronize()
    blk_ms = e0.elapsed_time(e1) / 90
    hbm = (peaks or {}).get("hbm_gbs", 6650.0)
    # algorithmic bytes per sample-block: bf16 x in + out (256 B); + the 80 sample-rate aux channels (160 B) only when
    # the aux projection cannot be taken at frame rate
    bytes_per = 256 if frames_op is not None else 416
    gbs = tracks * Tn * bytes_per / (blk_ms * 1e-3) / 1e9
    kernel = "usfgan_block_fr_kernel (frame-rate aux projection)" if frames_op is not None else "usfgan_block_kernel"
    return {"metric": "vocoded audio-sec/sec (ParallelHn-uSFGAN, 24 kHz)", "value": tracks * seconds / (ms / 1e3),
            "unit": "audio-sec/s", "ms_per_pass": ms, "precision": m.resolved_precision(),
            "config": {"workload": f"{tracks} tracks x {seconds:.0f} s @ 24 kHz, hop 120, aux 80, 20A+5F+30F blocks"},
            "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                         "traffic": _ncu_traffic(USFGAN_NCU_SUMMARY), "kernel": kernel, "us_per_launch": blk_ms * 1e3,
                         "bytes_per_sample_block": bytes_per,
                         "tflops": 2.0 * tracks * Tn * 38912 / (blk_ms * 1e-3) / 1e12,
                         "tensor_frac": 2.0 * tracks * Tn * 38912 / (blk_ms * 1e-3) / 1e12 / (peaks or {}).get("bf16_tflops_sustained", 1360.8)}}


def run_reference(args, rank):
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        v, kind, sample, dt, cores = reference_steps(torch.device("cpu"), 1)
        if i >= args.warmup:
            vals.append((v, dt))
    v = sum(x[0] for x in vals) / len(vals)
    line = {"metric": METRIC, "value": v,
            "unit": "frame-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(x[1] for x in vals) / len(vals), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": dict(CONFIG),
            "note": "each step = 1 DDPM step of the full batch (bounded sample of the 100-step pass) on the host cores",
            "cpu_baseline": {"value": v, "unit": "frame-steps/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "frame-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vocoder", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the config-4 pipeline (64 songs, strong scaling)")
    ap.add_argument("--pipeline-songs", type=int, default=64)
    ap.add_argument("--no-training", action="store_true", help="skip the config-5 DDP training step")
    ap.add_argument("--no-extras", action="store_true", help="skip wavenet / reference_gpu_eager (N = 1 only)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from ensemble_svs_with_interactions_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    m = build_model().to(dev)
    den = m.denoise_fn
    g = torch.Generator().manual_seed(1234 + rank)       # every rank: its own 6-track work item
    cond_host = torch.randn(B, T, CFG["encoder_hidden_dim"], generator=g).pin_memory()
    cond_dev = cond_host.to(dev)
    out_host = torch.empty(B, T, CFG["in_dim"]).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass_resident():
        return m.inference(cond_dev)

    def one_pass_e2e():
        c = cond_host.to(dev, non_blocking=True)
        y = m.inference(c)
        out_host.copy_(y, non_blocking=True)
        return y

    def timed(fn, n_warm, n_steps):
        for _ in range(n_warm):
            fn()
        barrier()
        total = 0.0
        n0 = _lib.launch_count
        for _ in range(n_steps):
            flush.fill_(1.0)                     # evict L2 between timed iterations (outside the event pair)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        n_launch = _lib.launch_count - n0
        barrier()
        t = torch.tensor([total], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / 1e3, n_launch

    with ClockSampler(local_rank) as clk:
        sec, gpu_launches = timed(one_pass_resident, args.warmup, args.steps)   # libsvsk launches, timed region only
        sec_e2e, _ = timed(one_pass_e2e, 1, args.steps)

    # dominant kernel: the residual stack.  When all its CTA pairs fit the device (they do at config 2) the 20 blocks of
    # one denoiser call are ONE launch (diffnet_stack_kernel); otherwise 20 launches of the per-layer kernel.  Timed
    # live: 10 denoiser calls' worth between two events on the launching stream.
    plan = den.bf16_plan()
    from ensemble_svs_with_interactions_b200 import ops
    condb, _ = ops.nct_to_ntc(cond_dev.transpose(1, 2).contiguous())
    table = m._step_table()
    sb = [tl[50] for tl in table]
    xb0 = torch.randn(B, T, plan.C, device=dev).to(torch.bfloat16)
    xb1 = torch.empty_like(xb0)
    xb2 = torch.empty_like(xb0)
    x32 = torch.randn(B, T, plan.C, device=dev)
    skip32 = torch.zeros(B, T, plan.C, device=dev)
    flags = torch.empty((B * 2 * ((T + 255) // 256),), device=dev, dtype=torch.int32)
    use_stack = (max(plan.dilations) <= 8 and os.environ.get("SVSK_DIFFNET_STACK", "1") != "0"
                 and ops.diffnet_stack_fits(B, T, plan.C, plan.H))

    # the launch as a sampling pass issues it: conditioner projection of all layers precomputed once per pass
    # (DiffNet.cond_projection_bf16), the kernel runs K = 3C per layer and adds the projection in its gating epilogue
    pcond = den.cond_projection_bf16(condb, plan) if use_stack else None

    def blocks():
        if use_stack:
            ops.diffnet_stack_bf16(xb0, xb1, xb2, skip32, condb, plan.w1p_all, plan.woutp_all, table[:, 50:51],
                                   plan.bout_all, flags, plan.dilations, stepbias_batch_stride=0,
                                   stepbias_layer_stride=table.stride(0), pcond=pcond)
            return
        cur, nxt = xb0, xb1
        for i, lw in enumerate(plan.layers):
            ops.diffnet_block_bf16(cur, nxt, x32, skip32, condb, lw["w1p"], lw["woutp"], sb[i], lw["bout"],
                                   dilation=lw["dilation"], stepbias_batch_stride=0, init_skip=(i == 0), write_x=True)
            cur, nxt = nxt, cur

    for _ in range(3):
        blocks()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        blocks()
    e1.record()
    e1.synchronize()
    launches_per_call = 1 if use_stack else len(plan.layers)
    block_ms = e0.elapsed_time(e1) / (reps * launches_per_call)
    flops_per_launch = 2.0 * B * T * BLOCK_MAC_PER_FRAME * (len(plan.layers) if use_stack else 1)
    # ALGORITHMIC work (655 360 MAC per frame and block, conditioner 1x1 included).  With the projection hoisted
    # out of the K-step loop the launch EXECUTES 2C*H MACs per frame and block less; both are reported.
    executed_flops = flops_per_launch
    if pcond is not None:
        executed_flops -= 2.0 * B * T * (2 * plan.C * plan.H) * len(plan.layers)
    peaks = _peaks()
    peak_tf = (peaks or {}).get("bf16_tflops_sustained", 1400.0)
    achieved_tf = flops_per_launch / (block_ms * 1e-3) / 1e12

    voc = vocoder_bench(dev, peaks) if not args.no_vocoder else None

    # BASELINE configs[3]: 64 songs of 6 tracks x 30 s through the FFConvLSTM encoders, both diffusion models, the device
    # post-processing and the vocoder, host to host; songs are split over the ranks by sharding.assign (strong scaling)
    pl = None
    if not args.no_pipeline:
        from types import SimpleNamespace
        from tools import bench_pipeline
        pl = bench_pipeline.run(SimpleNamespace(songs=args.pipeline_songs, warmup_songs=2, encoders=True, postprocess=True,
                                                breakdown=False), rank, world, dev)
    # BASELINE configs[4]: data-parallel training step, 6 x 1000 frames per rank, gradient all-reduce over NCCL
    tr = None
    if not args.no_training:
        from tools import bench_train
        tr = bench_train.run(10, 3, rank, local_rank, world, dev)

    if rank == 0:
        total_units = world * B * T * K_STEP * args.steps
        value = total_units / sec
        e2e = total_units / sec_e2e
        line = {
            "metric": METRIC,
            "value": value, "unit": "frame-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": dict(CONFIG),
            "precision": den.resolved_precision(), "audio_sec_per_sec_equiv": value / K_STEP / 200.0,
            "e2e": {"value": e2e, "unit": "frame-steps/s", "h2d_bytes_per_step": cond_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": 1e3 * sec_e2e / args.steps},
            "gpu_launches": gpu_launches,
            "clocks": clk.summary(),
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf,
                         "traffic": (_ncu_traffic(STACK_NCU_SUMMARY if pcond is not None else STACK_NCU_SUMMARY_IN_GEMM) if use_stack else None),
                         "kernel": ("diffnet_stack_kernel (all 20 residual blocks in one launch; CTA pairs, tcgen05 cta_group::2)"
                                    if use_stack else "diffnet_block3_kernel (one residual block; CTA pairs, tcgen05 cta_group::2)"),
                         "us_per_launch": block_ms * 1e3, "flops_per_launch": flops_per_launch,
                         "executed_flops_per_launch": executed_flops,
                         "executed_frac": executed_flops / (block_ms * 1e-3) / 1e12 / peak_tf,
                         "conditioner_projection": "hoisted (once per pass)" if pcond is not None else "in the GEMM",
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400"},
            "whole_pass_tflops": 2.0 * MAC_PER_FRAME_STEP * B * T * K_STEP * args.steps * world / sec / 1e12,
        }
        if voc is not None:
            voc["value"] *= world   # one 6-track vocoder batch per rank, no collective
            line["vocoder"] = voc
        if pl is not None:
            line["pipeline"] = {k: pl[k] for k in ("metric", "value", "unit", "songs", "scaling", "seconds", "ms_per_song_rank0",
                                                   "gpu_launches_rank0", "config")}
        if tr is not None:
            line["training"] = tr
        if world == 1 and not args.no_extras:
            line["wavenet"] = wavenet_bench(dev)
            # the reference's own modules in eager PyTorch on this GPU (cuDNN / cuBLAS, fp32 with torch's default TF32
            # convolution setting): the bar to beat on the same box
            v, kind, sample, dt, _ = reference_steps(dev, 2, repeats=3)
            line["reference_gpu_eager"] = {"value": v, "unit": "frame-steps/s", "kind": kind, "sample": sample,
                                           "ms_per_ddpm_step": 1e3 * dt / 2,
                                           "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32),
                                           "speedup_of_value": value / v}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(sample_steps=2)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
