"""CPU: sharding rule and the statistics all-reduce over gloo with world_size 2 (the N>1 host logic)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from v2v_b200 import dist as vd


def test_shard_indices_partition():
    for n in (0, 1, 7, 16, 10000):
        for w in (1, 2, 3, 8):
            parts = [vd.shard_indices(n, r, w) for r in range(w)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert [len(p) for p in parts] == vd.shard_counts(n, w)
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        vd.shard_indices(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = vd.shard_indices(n_clips, rank, world)
    # fake per-clip statistics that depend on the global clip index only
    stats = torch.tensor([[3 * i + 1, 2 * i] for i in mine], dtype=torch.int64).reshape(-1, 2)
    # the full vector of SURVEY §8(e): totals, per-bin |count| sums, per-pixel event-count map
    bins = torch.tensor([float(sum(mine)) * (b + 1) for b in range(5)], dtype=torch.float64)
    cmap = torch.full((3, 4), len(mine), dtype=torch.int64)
    vec = vd.pack_stats(stats, pixel_intervals=100 * len(mine), clips=len(mine), bin_abs_sums=bins, count_map=cmap)
    assert vec.numel() == 4 + 5 + 12
    vd.allreduce_stats(vec)
    u = vd.unpack_stats(vec, num_bins=5, map_shape=(3, 4))
    d = vd.stats_dict(vec)
    d["bin_abs_sums"], d["count_map"] = u["bin_abs_sums"].tolist(), u["count_map"].tolist()
    q.put((rank, d))
    dist.destroy_process_group()


def test_allreduce_stats_gloo_world2():
    world, n_clips = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = {"positive_events": sum(3 * i + 1 for i in range(n_clips)), "negative_events": sum(2 * i for i in range(n_clips)),
              "pixel_intervals": 100 * n_clips, "clips": n_clips,
              "bin_abs_sums": [sum(range(n_clips)) * (b + 1) for b in range(5)], "count_map": [[n_clips] * 4] * 3}
    assert all(d == expect for _, d in out)


def test_pack_stats_empty_rank():
    v = vd.pack_stats(None, 0, 0)
    assert v.tolist() == [0, 0, 0, 0]
    assert vd.allreduce_stats(v).tolist() == [0, 0, 0, 0]      # no process group: no-op
