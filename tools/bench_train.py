#!/usr/bin/env python
"""BASELINE configs[4]: data-parallel DiffNet training step (per rank: 6 x 1000-frame multi-track batch; 48 x 1000 on
8 GPUs), gradient all-reduce by DistributedDataParallel over NCCL — the reference's own scheme
(nnsvs/train_util.py:1444-1446, nnsvs/bin/train_acoustic_multitrack.py:358-380: L1 DDPM loss, clip_grad_norm_, AdamW).
Forward = libsvsk kernels; backward = interim autograd re-statement (see diffsinger/training.py).

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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(1234)
    model = GaussianDiffusion(256, 60, DiffNet(60, 256, 20, 256, 4), K_step=100)
    with torch.no_grad():
        model.denoise_fn.output_projection.weight.normal_(0, 0.02)
    model = model.to(dev).train()
    ddp = DDP(model, device_ids=[local])
    opt = torch.optim.AdamW(ddp.parameters(), lr=1e-3, betas=(0.9, 0.98))
    B, T = 6, 1000
    g = torch.Generator().manual_seed(1234 + rank)
    cond = torch.randn(B, T, 256, generator=g).to(dev)
    y = torch.randn(B, T, 60, generator=g).to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        noise, eps = ddp(cond, None, y)
        loss = (noise - eps).abs().mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ddp.parameters(), 10.0)
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record(); e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # every rank must hold identical parameters after the all-reduced updates
    chk = torch.stack([p.detach().float().sum() for p in model.parameters()]).sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = t.item() / args.steps
        n_par = sum(p.numel() for p in model.parameters())
        print(json.dumps({"metric": "DP DiffNet training step", "value": ms, "unit": "ms/step", "higher_is_better": False,
                          "n_gpus": world, "global_batch": f"{world * B} x {T} frames", "frames_per_sec": world * B * T / (ms / 1e3),
                          "loss": float(loss), "params_in_sync": bool(torch.equal(lo, hi)),
                          "allreduce_mb_per_step": n_par * 4 / 1e6,
                          "note": "forward: libsvsk bf16 kernels; backward: interim PyTorch autograd re-statement"}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
