"""world_size-2 gloo tests (CPU tensors) of the host-side exchange logic of the sharded build."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alga_b200 import readset
from alga_b200.distributed import interleave_shards, owner_of, route_triples


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, out_dir)
    finally:
        dist.destroy_process_group()


def _spawn(fn, tmp_path, world=2):
    mp.spawn(_worker, args=(world, _free_port(), fn, str(tmp_path)), nprocs=world, join=True)


def _all_triples(n_total, seed=5, m=5000):
    rng = np.random.default_rng(seed)
    return np.stack([rng.integers(0, n_total, m), rng.integers(0, n_total, m), rng.integers(0, 90, m)], 1).astype(np.int32)


def _route_fn(rank, world, out_dir):
    n_shard = 700
    allt = _all_triples(n_shard * world)
    mine = torch.from_numpy(allt[rank::world].copy())  # any initial distribution
    for col in (0, 1):
        got = route_triples(mine, col, n_shard, world)
        np.save(os.path.join(out_dir, f"route_{col}_{rank}.npy"), got.numpy())
    # empty send buffers on one rank
    got = route_triples(mine[:0] if rank == 0 else mine, 1, n_shard, world)
    np.save(os.path.join(out_dir, f"route_empty_{rank}.npy"), got.numpy())


def test_route_triples_gloo(tmp_path):
    world, n_shard = 2, 700
    _spawn(_route_fn, tmp_path, world)
    allt = _all_triples(n_shard * world)
    for col in (0, 1):
        for r in range(world):
            got = np.load(tmp_path / f"route_{col}_{r}.npy")
            want = allt[(allt[:, col] // n_shard) == r]
            assert got.shape == want.shape
            assert np.array_equal(got[np.lexsort(got.T[::-1])], want[np.lexsort(want.T[::-1])])
    sent = allt[1::world]
    for r in range(world):
        got = np.load(tmp_path / f"route_empty_{r}.npy")
        want = sent[(sent[:, 1] // n_shard) == r]
        assert np.array_equal(got[np.lexsort(got.T[::-1])], want[np.lexsort(want.T[::-1])])


def _interleave_fn(rank, world, out_dir):
    rng = np.random.default_rng(100 + rank)
    n = 2 * (50 + 3 * rank)  # chromosomes of different size
    rs = readset.from_code_matrix(rng.integers(0, 4, size=(n, 40), dtype=np.uint8))
    shard, total = interleave_shards(rs, rank, world, torch.device("cpu"))
    np.save(os.path.join(out_dir, f"shard_{rank}.npy"), shard.numpy())
    np.save(os.path.join(out_dir, f"src_{rank}.npy"), rs.words.view(np.int32).reshape(n, -1))
    assert total == shard.shape[0] * world


def test_interleave_shards_gloo(tmp_path):
    world = 2
    _spawn(_interleave_fn, tmp_path, world)
    src = [np.load(tmp_path / f"src_{r}.npy") for r in range(world)]
    twins = (min(s.shape[0] for s in src) // 2 // world) * world
    # global order: twin pair g = t * world + r  ->  rows (2g, 2g+1)
    glob = np.empty((2 * twins * world, src[0].shape[1]), np.int32)
    for r in range(world):
        for t in range(twins):
            g = t * world + r
            glob[2 * g: 2 * g + 2] = src[r][2 * t: 2 * t + 2]
    n_shard = 2 * twins
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"shard_{r}.npy"), glob[r * n_shard:(r + 1) * n_shard])


def test_owner_of_clamps():
    ids = torch.tensor([0, 9, 10, 19, 20, 25])
    assert owner_of(ids, 10, 2).tolist() == [0, 0, 1, 1, 1, 1]
