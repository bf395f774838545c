#!/usr/bin/env python
"""BASELINE configs[4]: data-parallel DiffNet training step (per rank: 6 x 1000-frame multi-track batch; 48 x 1000 on
8 GPUs), gradient all-reduce by DistributedDataParallel over NCCL — the reference's own scheme
(nnsvs/train_util.py:1444-1446, nnsvs/bin/train_acoustic_multitrack.py:358-380: L1 DDPM loss, clip_grad_norm_, AdamW).
Forward and backward = libsvsk tcgen05 kernels (svsk_seggemm_bf16 / svsk_wgrad_bf16, SURVEY §8(f) row 4, see
diffsinger/training.py), both captured in CUDA graphs.

  python tools/bench_train.py            |  torchrun --nproc-per-node N tools/bench_train.py
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: E402

from ensemble_svs_with_interactions_b200.diffsinger import DiffNet, GaussianDiffusion  # noqa: E402


def run(steps, warmup, rank, local, world, dev):
    """One measurement inside an already initialised process group (world > 1) or without one (world == 1).
    Returns the dict of the JSON line on rank 0, None elsewhere."""
    torch.manual_seed(1234)
    model = GaussianDiffusion(256, 60, DiffNet(60, 256, 20, 256, 4), K_step=100)
    with torch.no_grad():
        model.denoise_fn.output_projection.weight.normal_(0, 0.02)
    model = model.to(dev).train()
    # the whole stack's gradients appear at once (one autograd node): one bucket that holds them all, and .grad tensors that
    # ARE the bucket (no copy in, no copy out)
    ddp = DDP(model, device_ids=[local], gradient_as_bucket_view=True, bucket_cap_mb=128) if world > 1 else model
    # torch's multi-tensor ("fused") AdamW: the same optimizer the reference's trainer builds (a library call either way), one
    # kernel over the 171 parameter tensors instead of a dozen foreach launches (1.2 -> 0.3 ms of a 4.4 ms step)
    opt = torch.optim.AdamW(ddp.parameters(), lr=1e-3, betas=(0.9, 0.98), fused=True)
    B, T = 6, 1000
    g = torch.Generator().manual_seed(1234 + rank)
    cond = torch.randn(B, T, 256, generator=g).to(dev)
    y = torch.randn(B, T, 60, generator=g).to(dev)

    def step(sync=True):
        opt.zero_grad(set_to_none=True)
        if world > 1 and not sync:
            with ddp.no_sync():
                noise, eps = ddp(cond, None, y)
                loss = (noise - eps).abs().mean()
                loss.backward()
        else:
            noise, eps = ddp(cond, None, y)
            loss = (noise - eps).abs().mean()
            loss.backward()
        torch.nn.utils.clip_grad_norm_(ddp.parameters(), 10.0)
        opt.step()
        return loss

    def timed(sync):
        for _ in range(warmup):
            step(sync)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step(sync)
        e1.record(); e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / steps, loss

    # exposed communication = step time with the gradient all-reduce minus the same step without it (ranks drift apart
    # in the no_sync run, so it goes second and the parameters are checked after the synchronised one)
    if os.environ.get("SVSK_TRAIN_BREAKDOWN") and rank == 0:
        # per-phase times of one step, each bracketed by a synchronize (so they do not add up to the pipelined step time)
        import time
        for _ in range(3):
            step(True)
        ph = {}

        def lap(name, fn):
            torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
            ph[name] = ph.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
            return out
        for _ in range(5):
            lap("zero_grad", lambda: opt.zero_grad(set_to_none=True))
            noise, eps = lap("forward", lambda: ddp(cond, None, y))
            loss_ = lap("loss", lambda: (noise - eps).abs().mean())
            lap("backward", lambda: loss_.backward())
            lap("clip_grad_norm", lambda: torch.nn.utils.clip_grad_norm_(ddp.parameters(), 10.0))
            lap("optimizer", lambda: opt.step())
        print("breakdown ms/step:", {k: round(v / 5, 3) for k, v in ph.items()}, file=sys.stderr)
    ms, loss = timed(True)
    chk = torch.stack([p.detach().float().sum() for p in model.parameters()]).sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ms_nosync = timed(False)[0] if world > 1 else ms
    if rank != 0:
        return None
    n_par = sum(p.numel() for p in model.parameters())
    grad_bytes = n_par * 4
    exposed = max(ms - ms_nosync, 0.0)
    # ring all-reduce moves 2 (N-1)/N of the buffer per rank: bus bandwidth as NCCL's tests define it
    busbw = (2.0 * (world - 1) / world * grad_bytes / (exposed * 1e-3) / 1e9) if world > 1 and exposed > 0 else None
    from ensemble_svs_with_interactions_b200.diffsinger import training
    return {"metric": "DP DiffNet training step", "value": ms, "unit": "ms/step", "higher_is_better": False,
            "n_gpus": world, "global_batch": f"{world * B} x {T} frames", "frames_per_sec": world * B * T / (ms / 1e3),
            "tflops_per_gpu": 3 * 2 * 13_203_456 * B * T / (ms * 1e-3) / 1e12,
            "loss": float(loss), "params_in_sync": bool(torch.equal(lo, hi)),
            "allreduce_mb_per_step": grad_bytes / 1e6, "ms_per_step_without_allreduce": ms_nosync,
            "exposed_allreduce_ms": exposed, "allreduce_busbw_gbs_lower_bound": busbw,
            "backward": getattr(training, "BACKWARD_IMPL", "autograd re-statement"),
            "optimizer": "torch.optim.AdamW(fused=True) + clip_grad_norm_(10)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    line = run(args.steps, args.warmup, rank, local, world, dev)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
