"""Rank plumbing of the data-parallel learner (one process per GPU, torch.distributed).

Replaces the Horovod calls rl_games makes for the reference (phc/run_hydra.py:114-128; common_agent.py:112-127,
:225,:241; amp_agent.py:460-483,:650-673):
  * gradient averaging: ONE all-reduce(sum) over the flat gradient bucket per minibatch; the 1/world scale is applied
    inside the fused clip+Adam kernel (phc_adam_step's grad_scale) so every rank clips on the same reduced norm;
  * `hvd.sync_stats`: running mean/std statistics are averaged across ranks once per epoch;
  * per-rank seeds: seed + rank.
Works on any backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.distributed as dist


def is_multi() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size() -> int:
    return dist.get_world_size() if is_multi() else 1


def rank_seed(seed: int) -> int:
    return seed + (dist.get_rank() if is_multi() else 0)


def allreduce_grad_bucket(bucket: torch.Tensor) -> float:
    """Sum the flat gradient bucket over ranks in place; returns the scale (1/world) the optimiser must apply."""
    if not is_multi():
        return 1.0
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
    return 1.0 / dist.get_world_size()


class GradReducer:
    """The per-minibatch gradient all-reduce, issued OFF the compute stream: begin() makes a side stream wait for the backward
    pass, enqueues the collective there and returns at once (with the 1/world scale for the optimiser); end() makes the compute
    stream wait for it.  Whatever the caller enqueues on the compute stream in between overlaps the collective."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None

    def begin(self, bucket: torch.Tensor) -> float:
        if not is_multi():
            return 1.0
        if self.stream is None:                        # CPU tensors (gloo tests)
            dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
        else:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size()

    def end(self) -> None:
        if is_multi() and self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)


def broadcast_params(bucket: torch.Tensor, src: int = 0) -> None:
    if is_multi():
        dist.broadcast(bucket, src)


def sync_running_stats(stats: Iterable) -> None:
    """Average (running_mean, running_var, count) of every RunningMeanStd across ranks (rl_games' hvd.sync_stats)."""
    if not is_multi():
        return
    w = dist.get_world_size()
    for r in stats:
        if r is None:
            continue
        flat = torch.cat([r.running_mean.reshape(-1), r.running_var.reshape(-1), r.count.reshape(-1)])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= w
        n = r.running_mean.numel()
        r.running_mean.copy_(flat[:n].view_as(r.running_mean))
        r.running_var.copy_(flat[n:2 * n].view_as(r.running_var))
        r.count.copy_(flat[2 * n].view_as(r.count))


def max_over_ranks(value: float, device) -> float:
    if not is_multi():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
