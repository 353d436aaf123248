"""Multi-GPU plumbing: independent edits sharded over ranks, one model replica per GPU, no collective on the hot path
-- only a final result gather (reference: evaluation/FreeFine/freefine_batch_infer_2d.py:141-173, :242-262).

* `shard_indices(n, rank, world)` reproduces `DistributedSampler(dataset, shuffle=False, drop_last=False)` (:167): edit
  i goes to rank i mod world, and when n is not divisible the tail ranks are padded by wrapping around to the first
  samples (the reference then reports those edits twice; `gather_results` de-duplicates).
* `gather_results(local, indices)` all-gathers per-rank result tensors (+ their global indices) over
  torch.distributed -- NCCL over NVLink on the GPU box, gloo in the CPU tests -- and returns them ordered by index on
  every rank.  64 KiB of fp32 latents per 512x512 edit: latency-bound, nothing to fuse with.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def shard_indices(n: int, rank: int, world: int):
    """Global edit indices of `rank` (DistributedSampler, shuffle=False, drop_last=False)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    per = math.ceil(n / world) if n else 0
    total = per * world
    idx = list(range(n))
    if total > n and n > 0:
        pad = total - n
        idx += (idx * math.ceil(pad / n))[:pad]
    return idx[rank:total:world]


def gather_results(local: torch.Tensor, indices, n_total: int) -> torch.Tensor:
    """local [n_local, ...] results of the edits `indices` (as from shard_indices) -> [n_total, ...] on every rank,
    row i = result of edit i (first occurrence wins for the padded duplicates)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        seen = set()
        for row, i in zip(local, indices):
            if i not in seen:
                out[i] = row
                seen.add(i)
        return out
    world = dist.get_world_size()
    idx_t = torch.tensor(list(indices), dtype=torch.int64, device=local.device)
    counts = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(indices)], dtype=torch.int64, device=local.device))
    n_max = int(max(int(c) for c in counts))
    pad = lambda t: torch.cat([t, t.new_zeros((n_max - t.shape[0],) + tuple(t.shape[1:]))]) if t.shape[0] < n_max else t
    all_idx = [torch.empty(n_max, dtype=torch.int64, device=local.device) for _ in range(world)]
    all_res = [torch.empty((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(all_idx, pad(idx_t))
    dist.all_gather(all_res, pad(local.contiguous()))
    out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    seen = set()
    for r in range(world):
        for k in range(int(counts[r])):
            i = int(all_idx[r][k])
            if i not in seen:
                out[i] = all_res[r][k]
                seen.add(i)
    if len(seen) != n_total:
        raise RuntimeError(f"gathered {len(seen)} distinct edits, expected {n_total}")
    return out
