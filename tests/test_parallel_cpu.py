"""CPU tests of the N>1 host logic on the gloo backend (world_size 2): batch sharding is a partition, gather restores
image order (ragged shards included), and results do not depend on the world size."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pixart_sigma_b200.parallel import gather_batch, rank_seed, shard_batch, shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,world", [(32, 8), (5, 2), (3, 4), (1, 2), (7, 3)])
def test_shard_bounds_partition(n, world):
    spans = [shard_bounds(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank builds the same global problem, keeps only its shard, "denoises" it, gathers the result
        g = torch.Generator().manual_seed(0)
        x = torch.randn(n_items, 4, 8, 8, generator=g)
        y = torch.randn(n_items, 1, 6, 16, generator=g)
        xs, ys, none = shard_batch([x, y, None])
        assert none is None and xs.shape[0] == ys.shape[0]
        lo, hi = shard_bounds(n_items, rank, world)
        # stand-in for the per-image sampling loop: a deterministic function of the image and its sharding-independent seed
        local = torch.stack([xs[i] * 2 + ys[i].mean() + rank_seed(7, lo + i) % 5 for i in range(xs.shape[0])]) \
            if xs.shape[0] else xs.new_zeros((0, 4, 8, 8))
        full = gather_batch(local, n_items)
        want = torch.stack([x[i] * 2 + y[i].mean() + rank_seed(7, i) % 5 for i in range(n_items)])
        assert torch.equal(full, want)
        torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [4, 5])
def test_sharded_inference_gloo_world2(tmp_path, n_items):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_items, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(a, b) and a.shape[0] == n_items
