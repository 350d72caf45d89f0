"""CPU restatement of the reference's 2-hop graph construction, GraphStore.construct_coll_2hop
(code/graph_storage.py:127-246).  TEST INFRASTRUCTURE ONLY (tests/, tools/make_golden.py): the product path is
score_b200/csrc/hop2.cu.

The reference shuffles with Python's ``random`` (seeded 11, graph_storage.py:12) and caps the 2-hop lists with the
unseeded global NumPy generator (:187-189, :227-229), so its own output is not reproducible.  Here - and in the CUDA
kernel - the permutation of the k-th ``random.shuffle`` / ``np.random.choice`` call, calls counted in the reference's
processing order (all items, then all users; slices ascending), is the stable ascending argsort of the Philox4x32-10
uniforms ``philox_uniform(seed, 11 | 12, k, j)`` (``permutation`` below).

Pinned by tests/golden/hop2_reference.npz: documents produced by the reference's OWN method, executed unmodified from
/root/reference over an in-memory stand-in for MongoDB with ``random.shuffle`` and ``np.random.choice`` replaced by
functions that apply these permutations (tools/make_golden.py: make_hop2)."""
from __future__ import annotations

import numpy as np

from .loader_ref import philox_uniform

STREAM_SHUFFLE, STREAM_CHOICE = 11, 12


def permutation(seed, stream, call, n):
    """perm with new[r] = old[perm[r]]: stable ascending argsort of the call's n uniforms (csrc/hop2.cu)"""
    u = philox_uniform(seed, stream, call, np.arange(n, dtype=np.uint64))
    return np.argsort(u, kind="stable")


def build_2hop(user_1hop, item_1hop, n_user, n_item, n_slices, start_time, max_1hop, max_2hop, seed):
    """user_1hop[uid] / item_1hop[iid]: the per-slice 1-hop lists (construct_coll_1hop's documents, :90-125).
    Returns (user_docs, item_docs): id -> {'1hop', '2hop', 'degrees'} as construct_coll_2hop stores them; the input
    lists are copied, the copies carry the in-place shuffles (:171-173, :211-213)."""
    u1 = {u: [list(x) for x in user_1hop[u]] for u in range(1, n_user + 1)}
    i1 = {i: [list(x) for x in item_1hop[i]] for i in range(n_user + 1, n_user + n_item + 1)}
    calls = {"shuffle": 0, "choice": 0}

    def shuffle(lst):                      # random.shuffle(lst): in place
        perm = permutation(seed, STREAM_SHUFFLE, calls["shuffle"], len(lst))
        calls["shuffle"] += 1
        lst[:] = [lst[int(p)] for p in perm]

    def expand(own, t, nbr_lists):         # :168-192 (items) = :208-232 (users)
        ids2, deg2 = [], []
        nbrs = own[t]
        if len(nbrs) > max_1hop:
            shuffle(nbrs)
            nbrs = nbrs[:max_1hop]
        for n in nbrs:
            lst = nbr_lists[n][t]
            degree = len(lst)
            if 1 < degree <= max_1hop:
                ids2 += lst
                deg2 += [degree] * degree
            elif degree > max_1hop:
                ids2 += lst[:max_1hop]
                deg2 += [degree] * max_1hop
        if len(ids2) > max_2hop:
            idx = permutation(seed, STREAM_CHOICE, calls["choice"], len(ids2))
            calls["choice"] += 1
            ids2 = np.array(ids2)[idx].tolist()[:max_2hop]
            deg2 = np.array(deg2)[idx].tolist()[:max_2hop]
        return ids2, deg2

    def side(own_lists, nbr_lists, ids):
        docs = {}
        for node in ids:
            doc = {"1hop": own_lists[node], "2hop": [[] for _ in range(start_time)], "degrees": [[] for _ in range(start_time)]}
            for t in range(start_time, n_slices):
                a, b = expand(own_lists[node], t, nbr_lists)
                doc["2hop"].append(a)
                doc["degrees"].append(b)
            docs[node] = doc
        return docs

    item_docs = side(i1, u1, range(n_user + 1, n_user + n_item + 1))   # items first: they see the users' original order
    user_docs = side(u1, i1, range(1, n_user + 1))                      # users see the items' shuffled lists
    return user_docs, item_docs


def random_1hop(rng, n_user, n_item, n_slices, n_edges, hot_items=3, hot_share=0.35):
    """interaction lists with a few hot items / active users so that lists longer than max_1hop exist"""
    u1 = {u: [[] for _ in range(n_slices)] for u in range(1, n_user + 1)}
    i1 = {i: [[] for _ in range(n_slices)] for i in range(n_user + 1, n_user + n_item + 1)}
    for _ in range(n_edges):
        u = int(rng.integers(1, n_user + 1)) if rng.random() > 0.2 else int(rng.integers(1, min(n_user, 3) + 1))
        i = n_user + (int(rng.integers(1, n_item + 1)) if rng.random() > hot_share else int(rng.integers(1, hot_items + 1)))
        t = int(rng.integers(0, n_slices))
        u1[u][t].append(i)        # construct_coll_1hop appends in file order (:118-119); repeats are kept
        i1[i][t].append(u)
    return u1, i1
