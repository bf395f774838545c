#!/usr/bin/env python
"""BASELINE configs[3]: jaCappella_ritsu-shaped ensemble pipeline for N synthetic songs, track-sharded over the ranks.

Per song (6 voice parts x 30 s): mgc diffusion (M=60, H=256, C=256, L=20, K=100) + bap diffusion (M=5, C=H=128, L=10,
K=100) on pre-computed synthetic conditioning [6, 6000, H] (encoders / lf0 / vuv are outside the hot path),
then ParallelHn-uSFGAN with aux = 60 mgc + 5 bap (padded to 72) for 6 x 720 000 samples at 24 kHz.
Work item = song (its 6 tracks form one batch); items are assigned with sharding.assign; NO data-path collective.
Strong scaling: the total number of songs is fixed.

  python tools/bench_pipeline.py --songs 64            (1 GPU)
  torchrun --nproc-per-node N tools/bench_pipeline.py --songs 64
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ensemble_svs_with_interactions_b200 import _lib, postprocess, sharding  # noqa: E402
from ensemble_svs_with_interactions_b200.diffsinger import DiffNet, GaussianDiffusion  # noqa: E402
from ensemble_svs_with_interactions_b200.model import FFConvLSTM  # noqa: E402
from ensemble_svs_with_interactions_b200.pipeline import EnsembleSynthesizer  # noqa: E402
from ensemble_svs_with_interactions_b200.usfgan.models import ParallelHnUSFGANGenerator  # noqa: E402

FS, HOP, TRACKS, SECONDS = 24000, 120, 6, 30.0
FRAMES = int(SECONDS * 200)          # 5 ms acoustic frames
VFRAMES = int(SECONDS * FS / HOP)    # vocoder frames (hop 120 @ 24 kHz = 5 ms)


LING = 87   # linguistic features + lf0 per frame (recipe yaml in_dim)


def build(dev, encoders=False):
    torch.manual_seed(1234)
    if encoders:   # the recipe's FFConvLSTM encoders in front of both denoisers (yaml lines 104-116, 146-158)
        kw = dict(in_ph_start_idx=3, in_ph_end_idx=50, embed_dim=256, num_lstm_layers=2)
        enc_m = FFConvLSTM(LING, ff_hidden_dim=512, conv_hidden_dim=256, lstm_hidden_dim=128, out_dim=256, **kw)
        enc_b = FFConvLSTM(LING, ff_hidden_dim=256, conv_hidden_dim=128, lstm_hidden_dim=64, out_dim=128, **kw)
        mgc = GaussianDiffusion(LING, 60, DiffNet(60, 256, 20, 256, 4), encoder=enc_m, K_step=100)
        bap = GaussianDiffusion(LING, 5, DiffNet(5, 128, 10, 128, 4), encoder=enc_b, K_step=100)
    else:
        mgc = GaussianDiffusion(256, 60, DiffNet(60, 256, 20, 256, 4), K_step=100)
        bap = GaussianDiffusion(128, 5, DiffNet(5, 128, 10, 128, 4), K_step=100)
    for m in (mgc, bap):
        with torch.no_grad():
            m.denoise_fn.output_projection.weight.normal_(0, 0.02)
    pe = {"conv_layers": 3, "kernel_size": 5, "dilation": 1, "padding_mode": "replicate"}
    voc = ParallelHnUSFGANGenerator(periodicity_estimator_params=pe, aux_channels=65)
    with torch.no_grad():
        voc.periodicity_estimator.layers[-2].weight_v.normal_(0, 0.05)
    voc.remove_weight_norm()
    return mgc.to(dev).eval(), bap.to(dev).eval(), voc.to(dev).eval()


def song_inputs(song, dev, encoders=False):
    g = torch.Generator().manual_seed(1234 + song)
    if encoders:
        ling = torch.randn(TRACKS, FRAMES, LING, generator=g)
        ling[..., 3:50] = torch.nn.functional.one_hot(torch.randint(0, 47, (TRACKS, FRAMES), generator=g), 47).float()
        cond_mgc = cond_bap = ling
    else:
        cond_mgc = torch.randn(TRACKS, FRAMES, 256, generator=g)
        cond_bap = torch.randn(TRACKS, FRAMES, 128, generator=g)
    f0 = torch.empty(TRACKS, 1, VFRAMES).uniform_(110, 880, generator=g)
    d = (FS / (f0 * 4)).repeat_interleave(HOP, dim=-1)
    sig = torch.randn(TRACKS, 2, VFRAMES * HOP, generator=g) * 0.1
    return [t.pin_memory() for t in (cond_mgc, cond_bap, d, sig)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--songs", type=int, default=64)
    ap.add_argument("--warmup-songs", type=int, default=1)
    ap.add_argument("--encoders", action="store_true", help="run the recipe's FFConvLSTM encoders in front of both denoisers "
                    "(input = 87 linguistic features per frame instead of pre-computed conditioning)")
    ap.add_argument("--postprocess", action="store_true", help="GV post-filter of the mgc stream + 50 Hz trajectory smoothing of both "
                    "streams on the device between the diffusion models and the vocoder (gen.postprocess_acoustic's defaults)")
    ap.add_argument("--breakdown", action="store_true", help="also print ms per phase of the last song")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = run(args, rank, world, dev)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run(args, rank, world, dev):
    """One measurement (args: songs, warmup_songs, encoders, postprocess, breakdown); returns the JSON line's dict on rank 0."""
    mgc, bap, voc = build(dev, args.encoders)
    enc = EnsembleSynthesizer(mgc, bap, None)
    gv = torch.rand(60, device=dev) + 0.5
    mine = sharding.assign([FRAMES] * args.songs, world)[rank]
    host = song_inputs(0, dev, args.encoders)   # same shapes for every song; contents re-seeded per song below (cheap host RNG is not timed)

    phases = {}   # --breakdown: ms per phase of the last song (events around the three models)
    # Host <-> device traffic runs on its own stream, double-buffered: the inputs of song k+1 are uploaded and the
    # waveforms of song k-1 are downloaded (into pinned memory) while song k computes, and the host never blocks inside
    # the loop, so its ~700 kernel launches per song are issued ahead of the GPU.
    copy = torch.cuda.Stream(device=dev)
    dev_in = [None, None]      # uploaded inputs of the next song + the event that marks them complete
    pinned_out = [torch.empty((TRACKS, 1, VFRAMES * HOP), dtype=torch.float32).pin_memory() for _ in range(2)]
    out_done = [None, None]
    count = [0]

    def upload(slot):
        with torch.cuda.stream(copy):
            tensors = [t.to(dev, non_blocking=True) for t in host]
            e = torch.cuda.Event()
            e.record(copy)
        dev_in[slot] = (tensors, e)

    def synth(song):
        k = count[0]
        count[0] += 1
        main = torch.cuda.current_stream()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if args.breakdown else None
        if ev: ev[0].record()
        if dev_in[k & 1] is None:
            upload(k & 1)
        tensors, e_in = dev_in[k & 1]
        dev_in[k & 1] = None
        main.wait_event(e_in)
        for t in tensors:
            t.record_stream(main)
        cond_mgc, cond_bap, d, sig = tensors
        upload((k + 1) & 1)                               # next song's inputs, behind this song's compute
        if ev: ev[1].record()
        if args.encoders:                                 # both encoders side by side (pipeline.EnsembleSynthesizer._encode)
            cond_mgc, cond_bap = enc._encode(cond_mgc, cond_bap, [FRAMES] * TRACKS)
        m = mgc.inference(cond_mgc, cond_is_encoded=True)  # (6, 6000, 60)
        if ev: ev[2].record()
        b = bap.inference(cond_bap, cond_is_encoded=True)  # (6, 6000, 5)
        if ev: ev[3].record()
        if args.postprocess:
            m = postprocess.variance_scaling(gv, m, offset=2)
            mb = postprocess.lowpass_filter(torch.cat([m, b], dim=-1), 200, cutoff=50)
            m, b = mb[..., :60], mb[..., 60:]
        aux = torch.cat([m, b], dim=-1).transpose(1, 2)   # (6, 65, 6000) == vocoder frames at 5 ms
        aux = torch.nn.functional.pad(aux, (2, 2), mode="replicate").contiguous()
        wav = voc(sig, aux, d, wave_only=True)[0]
        if ev: ev[4].record()
        if out_done[k & 1] is not None:
            out_done[k & 1].synchronize()                 # the pinned buffer of two songs ago has long been written
        e_wav = torch.cuda.Event()
        e_wav.record(main)
        with torch.cuda.stream(copy):
            copy.wait_event(e_wav)
            pinned_out[k & 1].copy_(wav, non_blocking=True)   # D2H of the 6 waveforms
            out_done[k & 1] = torch.cuda.Event()
            out_done[k & 1].record(copy)
        wav.record_stream(copy)
        out = pinned_out[k & 1]
        if ev:
            ev[4].synchronize()
            names = ["h2d", "encoders + mgc diffusion" if args.encoders else "mgc diffusion", "bap diffusion",
                     "postprocess + vocoder" if args.postprocess else "vocoder"]
            phases.update({n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(names)})
        return out

    for s in range(args.warmup_songs):
        synth(s)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = _lib.launch_count
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in mine:
        synth(s)
    torch.cuda.current_stream().wait_stream(copy)         # the last downloads are part of the job
    e1.record()
    e1.synchronize()
    sec = sharding.max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)
    wall = time.perf_counter() - t0
    if rank != 0:
        return None
    audio = args.songs * TRACKS * SECONDS
    return {"metric": "ensemble pipeline audio-sec/sec (mgc+bap diffusion + ParallelHn-uSFGAN)",
            "value": audio / sec, "unit": "audio-sec/s", "n_gpus": world, "songs": args.songs,
            "seconds": sec, "wall_seconds_rank0": wall, "scaling": "strong",
            "ms_per_song_rank0": 1e3 * sec / max(1, len(mine)), "gpu_launches_rank0": _lib.launch_count - n0,
            "config": {"workload": f"{args.songs} songs x {TRACKS} tracks x {SECONDS:.0f} s, K=100, 24 kHz",
                       "encoders": bool(args.encoders), "postprocess": bool(args.postprocess)},
            **({"phases_ms_last_song": phases} if args.breakdown else {})}


if __name__ == "__main__":
    main()
