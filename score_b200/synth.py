"""Synthetic batches of the reference loader's shape.

The layout is the 8-tuple that ``GraphLoader`` yields (graph_loader.py:383) and that
``SCOREBASE.train`` consumes by index (score.py:103-110):
  [0] user_1hop [B,T,K,if]  [1] user_2hop [B,T,K,uf]  [2] item_1hop [B,T,K,uf]
  [3] item_2hop [B,T,K,if]  [4] target_user [B,uf]     [5] target_item [B,if]
  [6] label [B]             [7] length [B]
All ids index one table: 0 = dummy, 1..U users, U+1..U+I items, then side features.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Shape:
    """One row of the SURVEY.md section 8 shape table."""
    name: str
    feature_size: int
    eb_dim: int
    hidden_size: int
    max_time_len: int
    obj_per_time_slice: int
    user_fnum: int
    item_fnum: int
    n_user: int
    n_item: int
    batch: int
    length: int          # pred_time - start_time of the training split (graph_loader.py:382)

    def ctor_args(self):
        return (self.feature_size, self.eb_dim, self.hidden_size, self.max_time_len,
                self.obj_per_time_slice, self.user_fnum, self.item_fnum)

    @property
    def ids_per_sample(self):
        T, K = self.max_time_len, self.obj_per_time_slice
        return T * K * 2 * (self.user_fnum + self.item_fnum) + self.user_fnum + self.item_fnum


# train_score.py:23-54, 285-364 (T = TIME_SLICE_NUM - START_TIME - 1; train pred_time 9/6/38)
SHAPES = {
    "tmall": Shape("tmall", 1529672, 16, 32, 11, 10, 3, 4, 424170, 1090390, 100, 9),
    "tmall_t10": Shape("tmall_t10", 1529672, 16, 32, 10, 10, 3, 4, 424170, 1090390, 100, 8),
    "taobao": Shape("taobao", 1 + 984080 + 4049268 + 9405, 16, 32, 8, 10, 1, 2, 984080, 4049268, 1024, 6),
    "ccmr": Shape("ccmr", 1 + 4920695 + 190129 + 80172 + 213482 + 63 + 1044, 16, 32, 40, 10, 1, 5,
                  4920695, 190129, 1024, 38),
    "ccmr_k20": Shape("ccmr_k20", 1 + 4920695 + 190129 + 80172 + 213482 + 63 + 1044, 16, 32, 40, 20, 1, 5,
                      4920695, 190129, 1024, 38),
    "large_vocab": Shape("large_vocab", 200000001, 64, 128, 8, 10, 1, 1, 100000000, 100000000, 1024, 6),
    # one rank's shard of the large-vocab table at 8 GPUs (25 M rows, d=64: 19.2 GB of var+m+v) for 1-GPU kernel runs
    "large_vocab_shard": Shape("large_vocab_shard", 25000001, 64, 128, 8, 10, 1, 1, 12500000, 12500000, 1024, 6),
    # small shapes for tests / smoke
    "tiny": Shape("tiny", 5000, 16, 32, 6, 10, 3, 4, 1000, 3000, 24, 4),
    "tiny_tb": Shape("tiny_tb", 6000, 16, 32, 8, 10, 1, 2, 2000, 3500, 64, 6),
}


def _node_ids(rng, kind, shape: Shape, size, zipf):
    """Field 0 = node id, fields 1.. = side-feature ids above the item range."""
    fnum = shape.user_fnum if kind == "user" else shape.item_fnum
    lo, n = (1, shape.n_user) if kind == "user" else (shape.n_user + 1, shape.n_item)
    if zipf:
        r = rng.zipf(zipf, size=size)
        node = lo + (r - 1) % n
    else:
        node = rng.integers(lo, lo + n, size=size)
    out = np.empty(tuple(size) + (fnum,), np.int64)
    out[..., 0] = node
    feat_lo = 1 + shape.n_user + shape.n_item
    feat_n = max(shape.feature_size - feat_lo, 1)
    for f in range(1, fnum):
        if feat_lo >= shape.feature_size:   # no side-feature rows in this table
            out[..., f] = rng.integers(1, shape.feature_size, size=size)
        else:
            # side features are a deterministic function of the node in the real data; hash it
            out[..., f] = feat_lo + (node * (2654435761 + 40503 * f) + f) % feat_n
    return out


def make_batch(shape: Shape, batch=None, seed=1111, zipf=0.0, neg=1, dummy_frac=0.10,
               length=None, dtype=np.int32):
    """One batch; user side replicated over the ``1+neg`` consecutive samples of a user
    (graph_loader.py:360-364), labels 1,0,... (graph_loader.py:378-381), tail slices copy the
    last live slice (graph_loader.py:254-256), ``dummy_frac`` of live slices all-zero."""
    rng = np.random.default_rng(seed)
    B = batch or shape.batch
    T, K = shape.max_time_len, shape.obj_per_time_slice
    L = shape.length if length is None else length
    grp = 1 + neg
    n_user_rows = (B + grp - 1) // grp

    def side(kind, n, cyclic):
        ids = _node_ids(rng, kind, shape, (n, T, K), zipf)
        if cyclic:  # 1-hop: cyclic pad from a random true length in [1,K] (graph_loader.py:181-182)
            true_len = rng.integers(1, K + 1, size=(n, T))
            idx = np.arange(K)[None, None, :] % true_len[..., None]
            ids = np.take_along_axis(ids, idx[..., None].repeat(ids.shape[-1], -1), axis=2)
        dummy = rng.random((n, T)) < dummy_frac
        ids[dummy] = 0
        if L < T:
            ids[:, L:] = ids[:, L - 1:L]
        return ids

    u1 = side("item", n_user_rows, True)
    u2 = side("user", n_user_rows, False)
    rep = np.repeat(np.arange(n_user_rows), grp)[:B]
    u1, u2 = u1[rep], u2[rep]
    i1 = side("user", B, True)
    i2 = side("item", B, False)
    tu = _node_ids(rng, "user", shape, (n_user_rows,), zipf)[rep]
    ti = _node_ids(rng, "item", shape, (B,), zipf)
    label = (np.arange(B) % grp == 0).astype(dtype)
    length_arr = np.full(B, L, dtype)
    return tuple(np.ascontiguousarray(x.astype(dtype)) for x in (u1, u2, i1, i2, tu, ti)) + (label, length_arr)
