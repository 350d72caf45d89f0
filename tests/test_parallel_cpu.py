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


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_peer_push_offsets_land_rows_where_the_all_to_all_would(world):
    """The index arithmetic of the peer-memory exchange (csrc/shard.cu: send_off / recv_off from the count matrix), restated
    in NumPy for `world` in-process ranks, against the all-to-all semantics of ExchangePlan: a served row lands in the
    requester's staged table at the row its mini key points to, a gradient row lands in the owner's buffer at the element
    aligned with the owner's served list.  (The kernels themselves are checked on the GPU with `world` handles on one
    device, tests/test_parity_gpu.py::test_sharded_peer_push_emulated_ranks.)"""
    rng = np.random.default_rng(40 + world)
    V, d, n = 997, 4, 300
    table = rng.standard_normal((V, d)).astype(np.float32); table[0] = 0
    keys = [rng.integers(0, V, n + 7 * r).astype(np.int64) for r in range(world)]          # ragged position counts
    for k in keys:
        k[::5] = 0
    # per-rank plan (score_shard_plan): positions grouped by owner, stable; dummies last
    owner = [np.where(k != 0, k % world, world) for k in keys]
    order = [np.argsort(o, kind="stable") for o in owner]
    cm = np.stack([np.bincount(o, minlength=world + 1) for o in owner])                      # [world][world + 1]
    sel = [order[r][:cm[r, :world].sum()] for r in range(world)]
    send_rows = [keys[r][sel[r]] // world + 1 for r in range(world)]
    mini = []
    for r in range(world):
        m = np.zeros(len(keys[r]), np.int64)
        m[sel[r]] = np.arange(1, len(sel[r]) + 1)
        mini.append(m)
    send_off = np.concatenate([np.zeros((world, 1), np.int64), np.cumsum(cm[:, :world], 1)], 1)   # send_off(r, o)
    recv_off = np.concatenate([np.zeros((1, world), np.int64), np.cumsum(cm[:, :world], 0)], 0)   # recv_off(o, r) = [r][o]
    # ids all-to-all: owner o's `want` list = the ranks' chunks for o, in rank order
    want = [np.concatenate([send_rows[r][send_off[r, o]:send_off[r, o + 1]] for r in range(world)]) for o in range(world)]
    local = [np.zeros(((V + world - 1) // world + 1, d), np.float32) for _ in range(world)]
    for o in range(world):
        rows = np.arange(o, V, world)
        local[o][rows // world + 1] = table[rows]
        local[o][0] = 0
    staged = [np.zeros((len(keys[r]) + 1, d), np.float32) for r in range(world)]
    for me in range(world):                                   # shard_serve_push_kernel, element e of owner `me`
        for e in range(len(want[me])):
            r = int(np.searchsorted(recv_off[1:, me], e, side="right"))
            slot = send_off[r, me] + (e - recv_off[r, me]) + 1
            staged[r][slot] = local[me][want[me][e]]
    for r in range(world):
        assert np.array_equal(staged[r][mini[r]], table[keys[r]])
    grads = [rng.standard_normal((len(keys[r]), d)).astype(np.float32) for r in range(world)]
    owned = [np.full((len(want[o]), d), np.nan, np.float32) for o in range(world)]
    for me in range(world):                                   # shard_grad_push_kernel, send slot s of requester `me`
        for s in range(len(sel[me])):
            o = int(np.searchsorted(send_off[me, 1:], s, side="right"))
            dst = recv_off[me, o] + (s - send_off[me, o])
            owned[o][dst] = grads[me][sel[me][s]]
    acc = np.zeros((V, d), np.float64)
    for o in range(world):
        assert not np.isnan(owned[o]).any()
        glob = (want[o] - 1) * world + o                      # owner-local row -> global id
        np.add.at(acc, glob, owned[o].astype(np.float64))
    exp = np.zeros((V, d), np.float64)
    for r in range(world):
        nz = keys[r] != 0
        np.add.at(exp, keys[r][nz], grads[r][nz].astype(np.float64))
    assert np.allclose(acc, exp, atol=1e-12)
