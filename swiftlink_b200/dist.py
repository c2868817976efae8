"""Multi-GPU plumbing: one chain group (replicate) per rank, no traffic while sampling, and one
small all-reduce at the end to merge the LOD accumulators.

LODscores::merge_results (lod_score.h:98-105) log-sums the raw accumulators element-wise and adds
the counts.  Across ranks that is  max + ln(sum exp(s - max))  -- all_reduce(MAX), then
all_reduce(SUM) of exp(s - max) -- plus all_reduce(SUM) of the counts (about 400 KB at 10k SNPs x
5 positions: latency-bound on NVLink).  Works with the `nccl` backend on CUDA tensors and with
`gloo` on CPU tensors (the CPU tests use the latter with world_size 2).
"""
import math

import torch
import torch.distributed as dist

LOG_ZERO = -1.7976931348623157e308       # -DBL_MAX, the empty accumulator (cuda_common.cu:25)


def merge_lod(raw, count, group=None):
    """raw: float64 tensor of log-sum accumulators (LOG_ZERO = empty); count: int.
    Returns (merged raw, total count); every rank gets the result."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return raw, int(count)
    mx = raw.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    empty = mx <= LOG_ZERO
    ex = torch.where(raw <= LOG_ZERO, torch.zeros_like(raw), torch.exp(raw - torch.where(empty, torch.zeros_like(mx), mx)))
    dist.all_reduce(ex, op=dist.ReduceOp.SUM, group=group)
    merged = torch.where(empty, mx, mx + torch.log(ex))
    c = torch.tensor([int(count)], dtype=torch.int64, device=raw.device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return merged, int(c.item())


def merge_swap_stats(success, failure, group=None, device=None):
    """MC3 exchange statistics of every rank's ladder (mc3.cc:107-108,164-181): two small integer
    all-reduces over 2 x (chains - 1) counters.  Returns (success, failure) summed over ranks."""
    s = torch.as_tensor(success, dtype=torch.int64, device=device).clone()
    f = torch.as_tensor(failure, dtype=torch.int64, device=device).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(f, op=dist.ReduceOp.SUM, group=group)
    return s, f


def normalise(raw, count, trait_prob):
    """LODscores::get (lod_score.h:86-88)"""
    return (raw - math.log(max(int(count), 1)) - trait_prob) / math.log(10.0)


def chain_placement(n_replicates, world_size):
    """replicate r runs on rank r % world_size (one chain group per GPU)"""
    return [[r for r in range(n_replicates) if r % world_size == k] for k in range(world_size)]
