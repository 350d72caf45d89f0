"""world_size-2 gloo tests (CPU) of the host-side exchange logic behind the row-sharded / data-parallel step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from score_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        V, d, n = 101, 8, 400
        table = torch.from_numpy(np.random.default_rng(0).standard_normal((V, d)).astype(np.float32))
        table[0] = 0
        local = parallel.global_to_local_table(table, world, rank)
        assert local.shape[0] == parallel.shard_rows(V, world)
        rng = np.random.default_rng(10 + rank)
        keys = torch.from_numpy(rng.integers(0, V, size=n).astype(np.int32))
        keys[::7] = 0                                     # dummy / masked positions
        plan = parallel.ExchangePlan(keys, world)
        want = plan.exchange_ids()
        assert want.numel() == plan.n_recv and int(want.min()) >= 1
        served = local[want.long()]                       # owner-side gather
        staged = plan.return_rows(served)
        # forward: the staged mini-table addressed by mini_keys equals the global table addressed by keys
        assert torch.equal(staged[plan.mini_keys.long()], table[keys.long()])
        assert int((plan.mini_keys == 0).sum()) == int((keys == 0).sum())
        # backward: per-position gradient rows reach their owners aligned with `want`
        grads = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32))
        owned = plan.send_grads(grads)
        acc_local = torch.zeros_like(local, dtype=torch.float64)
        acc_local.index_add_(0, want.long(), owned.double())
        # reference: every rank's (keys, grads) scattered into a global table
        all_keys = [torch.empty_like(keys) for _ in range(world)]
        all_grads = [torch.empty_like(grads) for _ in range(world)]
        dist.all_gather(all_keys, keys)
        dist.all_gather(all_grads, grads)
        acc_global = torch.zeros(V, d, dtype=torch.float64)
        for k, g in zip(all_keys, all_grads):
            nz = k != 0
            acc_global.index_add_(0, k[nz].long(), g[nz].double())
        expect = parallel.global_to_local_table(acc_global, world, rank)
        assert torch.allclose(acc_local, expect, atol=1e-12)
        # deterministic order: grouped by source rank, ascending position inside
        src_sizes = plan.recv_counts
        assert sum(src_sizes) == want.numel()
        # the count matrix of the CUDA plan (one all-gather of [world + 1] counts per rank) gives the same split sizes as
        # the count all-to-all above, and buffer sizes every rank agrees on
        mine = torch.bincount(torch.where(keys != 0, keys.long() % world, torch.full_like(keys.long(), world)), minlength=world + 1)
        rows = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(rows, mine)
        cm = torch.stack(rows).reshape(-1).tolist()
        send_c, recv_c, max_recv, max_pos = parallel.split_count_matrix(cm, world, rank)
        assert send_c == plan.send_counts and recv_c == plan.recv_counts
        t = torch.tensor([plan.n_recv, n])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert [max_recv, max_pos] == t.tolist()
        # dense all-reduce + global loss composition
        g = torch.full((5,), float(rank + 1))
        dist.all_reduce(g)
        assert torch.equal(g, torch.full((5,), float(sum(range(1, world + 1)))))
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_exchange_plan_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / ("ok%d" % r)) for r in range(world))


def test_shard_layout_roundtrip():
    V, d, world = 23, 4, 4
    table = torch.arange(V * d, dtype=torch.float32).reshape(V, d)
    seen = torch.zeros(V, dtype=torch.bool)
    for r in range(world):
        loc = parallel.global_to_local_table(table, world, r)
        for v in range(r, V, world):
            assert torch.equal(loc[v // world + 1], table[v])
            seen[v] = True
        assert float(loc[0].abs().sum()) == 0.0
    assert bool(seen.all())


def test_exchange_capacity_is_a_common_multiple_of_1024():
    # the list capacity of the packed data-parallel exchange: >= every rank's count, multiple of 1024, never 0
    assert parallel.exchange_capacity([0, 0]) == 1024
    assert parallel.exchange_capacity([1, 1024]) == 1024
    assert parallel.exchange_capacity([1025, 7]) == 2048
    assert parallel.exchange_capacity([129259, 129022, 128870]) == 130048
    for counts in ([5], [1023, 1024, 1025], [10 ** 6]):
        cap = parallel.exchange_capacity(counts)
        assert cap % 1024 == 0 and cap >= max(counts) and cap - max(counts) < 1024 + (max(counts) == 0) * 1024
