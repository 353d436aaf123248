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
    row i = result of edit i (first occurrence wins for the padded duplicates).  Every rank holds ceil(n/world) edits
    (shard_indices pads the tail), so ONE all_gather_into_tensor moves the results and one moves the indices; the
    de-duplication is host-side index arithmetic on a single small D2H copy (no per-edit synchronisation)."""
    idx_local = torch.as_tensor(list(indices), dtype=torch.int64)
    if local.shape[0] != idx_local.numel():
        raise ValueError(f"{local.shape[0]} results for {idx_local.numel()} indices")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        all_idx, all_res = idx_local, local
    else:
        world = dist.get_world_size()
        n_loc = torch.tensor([idx_local.numel()], dtype=torch.int64, device=local.device)
        n_all = torch.empty(world, dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(n_all, n_loc)
        n_all = n_all.cpu()
        n_max = int(n_all.max())
        if int(n_all.min()) != n_max:       # not the sampler's layout: pad to the longest shard with index -1
            pad = n_max - idx_local.numel()
            idx_local = torch.cat([idx_local, torch.full((pad,), -1, dtype=torch.int64)])
            local = torch.cat([local, local.new_zeros((pad,) + tuple(local.shape[1:]))])
        all_idx_d = torch.empty(world * n_max, dtype=torch.int64, device=local.device)
        all_res = torch.empty((world * n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(all_idx_d, idx_local.to(local.device))
        dist.all_gather_into_tensor(all_res, local.contiguous())
        all_idx = all_idx_d.cpu()
    # first occurrence of every edit index, in rank-major order (what the reference's rank-0 merge keeps last is the same
    # value: a padded duplicate is the same edit computed twice)
    import numpy as np
    ai = all_idx.numpy()
    valid = np.nonzero(ai >= 0)[0]
    uniq, first = np.unique(ai[valid], return_index=True)
    if len(uniq) != n_total or (n_total and (uniq[0] != 0 or uniq[-1] != n_total - 1)):
        raise RuntimeError(f"gathered {len(uniq)} distinct edits, expected {n_total}")
    rows = torch.from_numpy(valid[first]).to(all_res.device)
    return all_res.index_select(0, rows)
