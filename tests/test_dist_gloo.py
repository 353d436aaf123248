"""N>1 path on CPU: world_size-2 gloo processes shard independent edits like the reference's
DistributedSampler(shuffle=False, drop_last=False) (freefine_batch_infer_2d.py:167) and gather the results."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from freefine_b200.dist import gather_results, shard_indices


def test_shard_indices_match_distributed_sampler():
    from torch.utils.data import DistributedSampler
    for n in (1, 7, 8, 9, 1024):
        for world in (1, 2, 4, 8):
            for rank in range(world):
                ref = list(DistributedSampler(range(n), num_replicas=world, rank=rank, shuffle=False, drop_last=False))
                assert shard_indices(n, rank, world) == ref, (n, world, rank)
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idx = shard_indices(n, rank, world)
        local = torch.stack([torch.full((4, 2, 2), float(i)) + rank * 0.0 for i in idx])   # "latents" of edit i
        out = gather_results(local, idx, n)
        ok = all(bool((out[i] == float(i)).all()) for i in range(n)) and out.shape == (n, 4, 2, 2)
        q.put((rank, ok, idx))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [6, 7])
def test_two_rank_gloo_shard_and_gather(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    by_rank = {r: idx for r, _, idx in res}
    assert sorted(set(by_rank[0] + by_rank[1])) == list(range(n))          # every edit is owned by some rank
    assert by_rank[0] == list(range(0, n + (n % 2), 2))[: len(by_rank[0])] or by_rank[0][0] == 0


def test_single_process_gather_dedupes():
    idx = shard_indices(5, 0, 1)
    out = gather_results(torch.arange(5.)[:, None], idx, 5)
    assert out.flatten().tolist() == [0., 1., 2., 3., 4.]
