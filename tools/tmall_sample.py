"""Host-side ETL for the reference's bundled Tmall sample (BASELINE.json config 1): raw interaction log -> unified
ids -> per-slice 1-hop / 2-hop neighbor lists -> target lines.  A deterministic restatement of the reference's offline
scripts, used only to build the small fixture tests/golden/tmall_sample.npz (tools/make_golden.py); the hot path never
runs it.

  feateng_tmall.py:30-134   id remapping (users, items, categories, sellers, brands, ages, genders share one id space
                            starting at 1), 15-day time slices from 2015-05-01, side-feature dicts
  graph_storage.py:90-246   1-hop lists in file order; 2-hop = union over <= 10 (shuffled) 1-hop neighbors of THEIR
                            1-hop lists in the same slice (each cut to 10, kept only if that degree > 1), capped at 100
  gen_target.py:79-121      one line per user active in slice `pred_time` with earlier history:
                            uid, first item of that slice, 99 negatives drawn uniformly with replacement

Differences, all forced: (1) ``user_info_format1.csv`` is not shipped (.MISSING_LARGE_BLOBS), so age / gender are
synthesised from the raw user id (8 age groups, 3 genders) - they only feed side-feature ids; (2) the reference remaps
ids in ``list(set)`` order, which depends on the process's string-hash seed; here ids are assigned in sorted order;
(3) the reference's shuffles / negative draws use the global ``random`` state; here a seeded generator.
"""
from __future__ import annotations

import datetime

import numpy as np

START_DAY = datetime.date(2015, 5, 1)     # feateng_tmall.py:10
SLICE_DAYS = 15                           # feateng_tmall.py:11
TIME_SLICE_NUM_STORE = 14                 # graph_storage.py:47
MAX_1HOP, MAX_2HOP = 10, 100              # graph_storage.py:43-44
NEG_SAMPLE_NUM = 99                       # gen_target.py


def parse_log(path):
    rows = []
    with open(path) as f:
        next(f)
        for line in f:
            uid, iid, cid, sid, bid, date, _ = line.strip().split(",")
            day = datetime.date(2015, int(date[:2]), int(date[2:]))
            rows.append((uid, iid, cid, sid, bid, (day - START_DAY).days // SLICE_DAYS))
    return rows


def build(path, seed=1111):
    rng = np.random.default_rng(seed)
    rows = [r for r in parse_log(path) if 0 <= r[5] < TIME_SLICE_NUM_STORE]
    ages = {u: "age%d" % (int(u) % 8) for u in {r[0] for r in rows}}
    genders = {u: "g%d" % ((int(u) // 8) % 3) for u in ages}
    remap, nxt = [], 1
    for values in ({r[0] for r in rows}, {r[1] for r in rows}, {r[2] for r in rows}, {r[3] for r in rows},
                   {r[4] for r in rows}, set(ages.values()), set(genders.values())):
        d = {}
        for v in sorted(values, key=lambda s: (len(s), s)):
            d[v] = nxt
            nxt += 1
        remap.append(d)
    um, im, cm, sm, bm, am, gm = remap
    n_user, n_item, feature_size = len(um), len(im), nxt
    S = TIME_SLICE_NUM_STORE
    user_feat = {str(um[u]): [am[ages[u]], gm[genders[u]]] for u in um}
    item_feat = {}
    u1 = {um[u]: [[] for _ in range(S)] for u in um}
    i1 = {im[i]: [[] for _ in range(S)] for i in im}
    for uid, iid, cid, sid, bid, t in rows:                     # graph_storage.py:114-117, file order
        item_feat[str(im[iid])] = [cm[cid], sm[sid], bm[bid]]
        u1[um[uid]][t].append(im[iid])
        i1[im[iid]][t].append(um[uid])

    def two_hop(own, other):
        out, deg = {}, {}
        for node in sorted(own):
            h2, dg = [], []
            for t in range(S):
                nbrs = list(own[node][t])
                if len(nbrs) > MAX_1HOP:                         # graph_storage.py:171-173
                    nbrs = [nbrs[j] for j in rng.permutation(len(nbrs))][:MAX_1HOP]
                ids, ds = [], []
                for nb in nbrs:
                    lst = other[nb][t]
                    d = len(lst)
                    if 1 < d <= MAX_1HOP:
                        ids += lst; ds += [d] * d
                    elif d > MAX_1HOP:
                        ids += lst[:MAX_1HOP]; ds += [d] * MAX_1HOP
                if len(ids) > MAX_2HOP:                          # graph_storage.py:186-189
                    idx = rng.permutation(len(ids))[:MAX_2HOP]
                    ids = [ids[j] for j in idx]; ds = [ds[j] for j in idx]
                h2.append(ids); dg.append(ds)
            out[node], deg[node] = h2, dg
        return out, deg
    i2, idg = two_hop(i1, u1)
    u2, udg = two_hop(u1, i1)
    user_docs = {u: {"uid": u, "1hop": u1[u], "2hop": u2[u], "degrees": udg[u]} for u in u1}
    item_docs = {i: {"iid": i, "1hop": i1[i], "2hop": i2[i], "degrees": idg[i]} for i in i1}
    targets = {}
    for pred in (9, 10, 11):                                     # train / validation / test (train_score.py:356-358)
        lines = []
        for u in sorted(u1):
            if u1[u][pred] and any(u1[u][t] for t in range(pred)):      # gen_target.py:108-111
                negs = rng.integers(n_user + 1, n_user + n_item + 1, NEG_SAMPLE_NUM).tolist()
                lines.append([u, u1[u][pred][0]] + negs)
        targets[pred] = np.asarray(lines, np.int32)
    return dict(n_user=n_user, n_item=n_item, feature_size=feature_size, n_slices=S, user_docs=user_docs,
                item_docs=item_docs, user_feat=user_feat, item_feat=item_feat, targets=targets)
