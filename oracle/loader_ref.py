"""CPU restatement of the reference loader: GraphHandler neighbor selection (code/score/graph_loader.py:94-277) and
GraphLoader.worker batch assembly (:340-385).  TEST INFRASTRUCTURE ONLY (tests/, tools/make_golden.py): the product
path is score_b200/csrc/sampler.cu.

The reference draws 2-hop neighbors with ``np.random.choice`` (global NumPy state, not reproducible).  Here the K
uniforms of a (side, entity, slice) draw come from a caller-supplied function - by default the Philox4x32-10 stream
the CUDA sampler uses, restated in NumPy below - and are turned into indices the way NumPy does it
('rs': scaled uniform; 'is': ``cdf.searchsorted(u, side='right')`` on the normalised cumulative sum).

Pinned by tests/golden/loader_reference.npz: outputs of the reference's OWN GraphHandler methods (executed unmodified
from /root/reference with an in-memory stand-in for MongoDB and np.random.choice fed the same uniforms)."""
from __future__ import annotations

import copy

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox_uniform(seed, stream_id, step, idx):
    """uniform [0,1) of element `idx` (array ok) of stream (`stream_id`, `step`): csrc/common.cuh philox_uniform"""
    idx = np.asarray(idx, np.uint64)
    c = [(idx >> np.uint64(2)) & np.uint64(MASK), (idx >> np.uint64(34)) & np.uint64(MASK),
         np.full(idx.shape, stream_id, np.uint64), np.full(idx.shape, step, np.uint64)]
    k0, k1 = np.uint64(seed & MASK), np.uint64((seed >> 32) & MASK)
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & np.uint64(MASK)
        k1 = (k1 + np.uint64(W1)) & np.uint64(MASK)
    sel = (idx & np.uint64(3)).astype(np.int64)
    w = np.choose(sel, c)
    return ((w >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def draw_uniforms(seed, draw_id, side, ent, ts, T, K):
    """the K uniforms of one 2-hop draw; side 1 = user histories, 2 = item histories (sampler.cu key)"""
    key = (np.uint64(ent) * np.uint64(T) + np.uint64(ts)) * np.uint64(K) + np.arange(K, dtype=np.uint64)
    return philox_uniform(seed, side, draw_id, key)


def choose_rs(lst, u):
    n = len(lst)
    j = np.minimum((u.astype(np.float32) * np.float32(n)).astype(np.int64), n - 1)
    return [lst[int(x)] for x in j]


def choose_is(lst, degrees, u):
    """np.random.choice(lst, K, p=softmax(1/(deg-1))) given the uniforms (graph_loader.py:112-114)"""
    p = 1 / (np.array(degrees) - 1)
    p = np.exp(p) / np.sum(np.exp(p))
    cdf = p.cumsum()
    cdf /= cdf[-1]
    j = np.minimum(cdf.searchsorted(u.astype(np.float64), side="right"), len(lst) - 1)
    return [lst[int(x)] for x in j]


class GraphHandlerRef(object):
    """graph_loader.py:40-277 over in-memory documents ({uid: doc}, {iid: doc})."""

    def __init__(self, time_slice_num, user_docs, item_docs, obj_per_time_slice, user_num, item_num, start_time, mode,
                 user_feat_dict, item_feat_dict, user_fnum, item_fnum, uniforms):
        self.mode = mode
        self.user_docs, self.item_docs = user_docs, item_docs
        self.user_num, self.item_num, self.start_time = user_num, item_num, start_time
        self.obj_per_time_slice, self.time_slice_num = obj_per_time_slice, time_slice_num
        self.user_feat_dict, self.item_feat_dict = user_feat_dict, item_feat_dict
        self.user_fnum, self.item_fnum = user_fnum, item_fnum
        self.user_dummy_node = np.zeros([obj_per_time_slice, user_fnum]).tolist()     # float zeros, :90-91
        self.item_dummy_node = np.zeros([obj_per_time_slice, item_fnum]).tolist()
        self.uniforms = uniforms            # (side, ent, ts) -> K uniforms

    def _neighbors(self, doc, time_slice, side, ent, ts, fnum_1hop, feat_1hop, dummy_1hop, fnum_2hop, feat_2hop, dummy_2hop):
        K = self.obj_per_time_slice
        node_1hop_list = list(doc['1hop'][time_slice])
        node_2hop_list = doc['2hop'][time_slice]
        degree_list = doc['degrees'][time_slice]
        if node_1hop_list != []:
            if len(node_1hop_list) > K:
                node_1hop_list = node_1hop_list[:K]
            else:
                for i in range(K - len(node_1hop_list)):
                    node_1hop_list.append(node_1hop_list[i % len(node_1hop_list)])
            result_1hop = [[nid] if fnum_1hop == 1 else [nid] + feat_1hop[str(nid)] for nid in node_1hop_list]
        else:
            result_1hop = dummy_1hop
        if node_2hop_list != []:
            u = self.uniforms(side, ent, ts)
            picked = choose_is(node_2hop_list, degree_list, u) if self.mode == 'is' else choose_rs(node_2hop_list, u)
            result_2hop = [[nid] if fnum_2hop == 1 else [nid] + feat_2hop[str(nid)] for nid in picked]
        else:
            result_2hop = dummy_2hop
        return result_1hop, result_2hop

    def gen_user_history(self, start_uid, pred_time, ent):
        doc = self.user_docs[start_uid]
        user_1hop, user_2hop = [], []
        for i in range(self.start_time, pred_time):
            a, b = self._neighbors(doc, i, 1, ent, i - self.start_time, self.item_fnum, self.item_feat_dict,
                                   self.item_dummy_node, self.user_fnum, self.user_feat_dict, self.user_dummy_node)
            user_1hop.append(a); user_2hop.append(b)
        for i in range(self.time_slice_num - pred_time - 1):
            user_1hop.append(user_1hop[-1]); user_2hop.append(user_2hop[-1])
        return user_1hop, user_2hop

    def gen_item_history(self, start_iid, pred_time, ent):
        doc = self.item_docs[start_iid]
        item_1hop, item_2hop = [], []
        for i in range(self.start_time, pred_time):
            a, b = self._neighbors(doc, i, 2, ent, i - self.start_time, self.user_fnum, self.user_feat_dict,
                                   self.user_dummy_node, self.item_fnum, self.item_feat_dict, self.item_dummy_node)
            item_1hop.append(a); item_2hop.append(b)
        for i in range(self.time_slice_num - pred_time - 1):
            item_1hop.append(item_1hop[-1]); item_2hop.append(item_2hop[-1])
        return item_1hop, item_2hop


def assemble_batch(handler, uids, iids, pred_time, start_time, neg_sample_num):
    """GraphLoader.worker (graph_loader.py:340-385); entity numbering as in sampler.cu: user group e -> e, sample b ->
    n_groups + b.  Returns the 8-tuple of nested lists."""
    grp = neg_sample_num + 1
    n_groups = len(uids)
    u1b, u2b, i1b, i2b, tub, tib, lab, ln = [], [], [], [], [], [], [], []
    for i in range(len(uids)):
        user_1hop, user_2hop = handler.gen_user_history(uids[i], pred_time, i)
        for j in range(i * grp, (i + 1) * grp):
            item_1hop, item_2hop = handler.gen_item_history(iids[j], pred_time, n_groups + j)
            u1b.append(user_1hop); u2b.append(user_2hop); i1b.append(item_1hop); i2b.append(item_2hop)
            tub.append([uids[i]] if handler.user_feat_dict is None else [uids[i]] + handler.user_feat_dict[str(uids[i])])
            tib.append([iids[j]] if handler.item_feat_dict is None else [iids[j]] + handler.item_feat_dict[str(iids[j])])
            lab.append(1 if j % grp == 0 else 0)
            ln.append(pred_time - start_time)
    return (u1b, u2b, i1b, i2b, tub, tib, lab, ln)


def random_graph(rng, n_user, n_item, n_slices, max_deg=14, empty_frac=0.3, user_fnum=1, item_fnum=1, n_feat=50):
    """documents of the reference's shape for tests: per slice a (possibly empty, possibly > K) 1-hop list and a 2-hop
    list with degrees >= 2; feature dicts keyed by str(id) as the pickles are"""
    user_docs, item_docs = {}, {}
    for uid in range(1, n_user + 1):
        d = {'uid': uid, '1hop': [], '2hop': [], 'degrees': []}
        for _ in range(n_slices):
            n1 = 0 if rng.random() < empty_frac else int(rng.integers(1, max_deg + 1))
            n2 = 0 if rng.random() < empty_frac else int(rng.integers(1, 3 * max_deg))
            d['1hop'].append(rng.integers(n_user + 1, n_user + n_item + 1, n1).tolist())
            d['2hop'].append(rng.integers(1, n_user + 1, n2).tolist())
            d['degrees'].append(rng.integers(2, 9, n2).tolist())
        user_docs[uid] = d
    for iid in range(n_user + 1, n_user + n_item + 1):
        d = {'iid': iid, '1hop': [], '2hop': [], 'degrees': []}
        for _ in range(n_slices):
            n1 = 0 if rng.random() < empty_frac else int(rng.integers(1, max_deg + 1))
            n2 = 0 if rng.random() < empty_frac else int(rng.integers(1, 3 * max_deg))
            d['1hop'].append(rng.integers(1, n_user + 1, n1).tolist())
            d['2hop'].append(rng.integers(n_user + 1, n_user + n_item + 1, n2).tolist())
            d['degrees'].append(rng.integers(2, 9, n2).tolist())
        item_docs[iid] = d
    feat_lo = n_user + n_item + 1
    ufd = None if user_fnum == 1 else {str(u): rng.integers(feat_lo, feat_lo + n_feat, user_fnum - 1).tolist()
                                       for u in range(1, n_user + 1)}
    ifd = None if item_fnum == 1 else {str(i): rng.integers(feat_lo, feat_lo + n_feat, item_fnum - 1).tolist()
                                       for i in range(n_user + 1, n_user + n_item + 1)}
    return user_docs, item_docs, ufd, ifd


def deep(docs):
    return copy.deepcopy(docs)
