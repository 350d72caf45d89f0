#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ (run in the BUILD container, never on the GPU box).

Two kinds of fixture, and each file says which it is in its ``source`` field:

* ``metrics_reference.npz`` - outputs of the REFERENCE'S OWN ranking-metric functions
  (code/score/train_score.py:104-142: getNDCG_at_K, getHR_at_K, getMRR, get_ranking_quality).  train_score.py
  cannot be imported (it imports tensorflow at module level), so the function definitions are lifted out of the
  file with ``ast`` and executed unmodified against NumPy; log-loss / AUC come from scikit-learn as the reference
  calls them (train_score.py:158-159).  These pin oracle/metrics_ref.py and the CUDA metrics kernel to the
  reference itself.
* ``model_<case>.npz`` - outputs of the CPU restatement oracle/score_ref.py (TensorFlow 1.x cannot be installed
  here, so the model arithmetic has no reference-produced vectors: "parity unpinned").  They
  freeze the oracle (any later edit of the restatement shows up as a diff) and let the GPU box check the CUDA
  path without executing the oracle.

Usage:  python tools/make_golden.py [--reference /root/reference]
"""
from __future__ import annotations

import argparse
import ast
import hashlib
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

METRIC_FUNCS = ("getNDCG_at_K", "getHR_at_K", "getMRR", "get_ranking_quality")


def load_reference_metric_functions(reference_root):
    """exec the reference's metric function definitions (source text untouched) in a private namespace"""
    path = os.path.join(reference_root, "code", "score", "train_score.py")
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np, "math": math, "TEST_NEG_SAMPLE_NUM": 99}
    for node in tree.body:
        if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "TEST_NEG_SAMPLE_NUM" for t in node.targets):
            exec(compile(ast.Module([node], []), path, "exec"), ns)
        if isinstance(node, ast.FunctionDef) and node.name in METRIC_FUNCS:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


def metric_cases():
    """(name, preds, iids) covering tie-free groups, saturated ties, positives at every rank region, duplicate ids"""
    rng = np.random.default_rng(20240917)
    G = 100
    cases = []
    p = rng.random(7 * G).astype(np.float32)                     # tie-free
    iid = rng.permutation(100000)[:7 * G].astype(np.int32) + 1
    cases.append(("tie_free", p, iid))
    p = rng.random(5 * G).astype(np.float32)
    for g in range(5):                                           # the positive placed at chosen ranks
        order = np.argsort(-p[g * G:(g + 1) * G])
        want = [0, 4, 5, 9, 57][g]
        j = order[want]
        p[g * G], p[g * G + j] = p[g * G + j], p[g * G]
    iid = rng.permutation(100000)[:5 * G].astype(np.int32) + 1
    cases.append(("positive_at_ranks", p, iid))
    p = np.round(rng.random(6 * G), 1).astype(np.float32)        # heavy ties, positives untied
    p[::G] = (0.05 + 0.1 * np.arange(6)).astype(np.float32)
    iid = rng.permutation(100000)[:6 * G].astype(np.int32) + 1
    cases.append(("ties_not_on_positive", p, iid))
    p = rng.random(4 * G).astype(np.float32)                     # duplicate candidate ids among the negatives
    iid = rng.integers(1, 60, 4 * G).astype(np.int32) + 1000
    iid[::G] = np.arange(4) + 1                                  # positives unique
    cases.append(("duplicate_negative_ids", p, iid))
    return cases


def make_metrics(reference_root):
    from sklearn.metrics import log_loss, roc_auc_score
    ns = load_reference_metric_functions(reference_root)
    out = {"source": np.array("reference: code/score/train_score.py:104-142 executed unmodified (ast-extracted) + "
                              "sklearn log_loss/roc_auc_score as called at train_score.py:158-159")}
    names = []
    for name, p, iid in metric_cases():
        G = ns["TEST_NEG_SAMPLE_NUM"] + 1
        labels = (np.arange(len(p)) % G == 0).astype(np.int32)
        rq = ns["get_ranking_quality"](p.tolist(), iid.tolist())
        pl = [float(x) for x in p]
        res = np.array([log_loss(labels.tolist(), pl), roc_auc_score(labels.tolist(), pl)] + [float(x) for x in rq], np.float64)
        out[name + "/preds"], out[name + "/iids"], out[name + "/labels"], out[name + "/expect"] = p, iid, labels, res
        names.append(name)
    # the scalar helpers on explicit rank lists
    rl = list(range(10, 30))
    out["helpers/ndcg"] = np.array([ns["getNDCG_at_K"](rl, t, k) for t in (10, 12, 14, 15, 99) for k in (5, 10)], np.float64)
    out["helpers/hr"] = np.array([ns["getHR_at_K"](rl, t, k) for t in (10, 12, 14, 15, 99) for k in (1, 5, 10)], np.float64)
    out["helpers/mrr"] = np.array([ns["getMRR"](rl, t) for t in (10, 12, 14, 15, 29, 99)], np.float64)
    out["cases"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "metrics_reference.npz"), **out)
    print("metrics_reference.npz:", names)


# ------------------------------------------------------------------------------------------ loader (graph_loader.py)
class _FakeColl(object):
    def __init__(self, docs, key):
        self.docs, self.key = docs, key

    def find(self, q):
        import copy
        return [copy.deepcopy(self.docs[q[self.key]])]   # a Mongo query returns a fresh document every time


class _FakeDB(object):
    def __init__(self, user_docs, item_docs):
        self.user_docs, self.item_docs = user_docs, item_docs

    def __getitem__(self, name):
        return _FakeColl(self.user_docs, 'uid') if name.startswith('user_') else _FakeColl(self.item_docs, 'iid')


class _FakeMongo(object):
    def __init__(self, user_docs, item_docs):
        self.db = _FakeDB(user_docs, item_docs)

    def MongoClient(self, url):
        return {None: self.db}.get(None) and type("C", (), {"__getitem__": lambda s_, n: self.db})()


class _Ctx(object):
    side = 1; ent = 0; queue = None; seed = 0; draw_id = 0; T = 0; K = 0


def _np_proxy(ctx):
    """numpy with np.random.choice replaced: the K uniforms come from the Philox stream of the CUDA sampler (keyed by
    the draw the harness announced), the index rule is NumPy's own"""
    from oracle import loader_ref as L

    def choice(a, size=None, replace=True, p=None):
        ts = ctx.queue.pop(0)
        u = L.draw_uniforms(ctx.seed, ctx.draw_id, ctx.side, ctx.ent, ts, ctx.T, ctx.K)[:size]
        a = np.asarray(a)
        if p is None:
            idx = np.minimum((u * np.float32(len(a))).astype(np.int64), len(a) - 1)   # stands in for randint(0, n)
        else:
            cdf = np.asarray(p, np.float64).cumsum()
            cdf /= cdf[-1]
            idx = np.minimum(cdf.searchsorted(u.astype(np.float64), side='right'), len(a) - 1)   # numpy/random/mtrand.pyx choice()
        return a[idx]

    class _Random(object):
        pass
    rnd = _Random()
    rnd.choice = choice

    class _NP(object):
        random = rnd

        def __getattr__(self, name):
            return getattr(np, name)
    return _NP()


def load_reference_graph_handler(reference_root, ns):
    path = os.path.join(reference_root, "code", "score", "graph_loader.py")
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "GraphHandler":
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns["GraphHandler"]


LOADER_CASES = [
    # name, n_user, n_item, time_slice_num, start, pred, K, uf, if, mode, neg, n_groups
    ("rs_plain", 30, 40, 9, 0, 6, 10, 1, 1, "rs", 1, 6),
    ("rs_tmall_fields", 25, 35, 12, 0, 9, 10, 3, 4, "rs", 1, 5),
    ("rs_start2_neg3", 20, 30, 10, 2, 7, 5, 1, 2, "rs", 3, 4),
    ("is_taobao_fields", 30, 40, 9, 0, 7, 10, 1, 2, "is", 1, 6),
    ("rs_eval_last_slice", 20, 30, 9, 0, 8, 10, 1, 5, "rs", 9, 2),
]


def make_loader(reference_root):
    from oracle import loader_ref as L
    from score_b200.graph import docs_to_csr, feat_table
    out = {"source": np.array("reference: GraphHandler.gen_user_history / gen_item_history of code/score/graph_loader.py:"
                              "40-277 executed unmodified (ast-extracted class; MongoDB replaced by in-memory documents; "
                              "np.random.choice fed the Philox uniforms of csrc/sampler.cu), batch assembled in the order "
                              "of GraphLoader.worker :340-385")}
    names = []
    for ci, (name, nu, ni, tsn, start, pred, K, uf, fi, mode, neg, n_groups) in enumerate(LOADER_CASES):
        rng = np.random.default_rng(500 + ci)
        user_docs, item_docs, ufd, ifd = L.random_graph(rng, nu, ni, tsn, user_fnum=uf, item_fnum=fi)
        T = tsn - start - 1
        ctx = _Ctx()
        ctx.seed, ctx.draw_id, ctx.T, ctx.K = 777 + ci, ci, T, K
        ns = {"pymongo": _FakeMongo(user_docs, item_docs), "np": _np_proxy(ctx), "pkl": None}
        GH = load_reference_graph_handler(reference_root, ns)
        gh = GH(tsn, "db", K, nu, ni, start, 7, 11, mode, None, None, uf, fi)
        gh.user_feat_dict, gh.item_feat_dict = ufd, ifd
        grp = neg + 1
        uids = rng.integers(1, nu + 1, n_groups).tolist()
        iids = rng.integers(nu + 1, nu + ni + 1, n_groups * grp).tolist()
        cols = [[] for _ in range(8)]
        for i, uid in enumerate(uids):      # GraphLoader.worker, graph_loader.py:357-382
            ctx.side, ctx.ent = 1, i
            ctx.queue = [t for t in range(pred - start) if user_docs[uid]['2hop'][start + t] != []]
            u1, u2 = gh.gen_user_history(uid, pred)
            for j in range(i * grp, (i + 1) * grp):
                ctx.side, ctx.ent = 2, n_groups + j
                ctx.queue = [t for t in range(pred - start) if item_docs[iids[j]]['2hop'][start + t] != []]
                i1, i2 = gh.gen_item_history(iids[j], pred)
                cols[0].append(u1); cols[1].append(u2); cols[2].append(i1); cols[3].append(i2)
                cols[4].append([uid] if ufd is None else [uid] + ufd[str(uid)])
                cols[5].append([iids[j]] if ifd is None else [iids[j]] + ifd[str(iids[j])])
                cols[6].append(1 if j % grp == 0 else 0)
                cols[7].append(pred - start)
        off1, ids1, off2, ids2, deg2 = docs_to_csr(user_docs, item_docs, nu, ni, tsn)
        pre = name + "/"
        out[pre + "params"] = np.array([nu, ni, tsn, start, pred, K, uf, fi, 0 if mode == "rs" else 1, neg, 777 + ci, ci], np.int64)
        out[pre + "hop1_off"], out[pre + "hop1_ids"], out[pre + "hop2_off"], out[pre + "hop2_ids"], out[pre + "hop2_deg"] = off1, ids1, off2, ids2, deg2
        out[pre + "user_feat"] = feat_table(ufd, 1, nu, uf - 1)
        out[pre + "item_feat"] = feat_table(ifd, nu + 1, ni, fi - 1)
        out[pre + "uids"], out[pre + "iids"] = np.asarray(uids, np.int32), np.asarray(iids, np.int32)
        for k in range(8):
            out[pre + "batch/%d" % k] = np.asarray(cols[k]).astype(np.int32)    # feeding int32 placeholders casts (score.py:21-30)
        names.append(name)
        print("loader case %s: B=%d T=%d, %d 2-hop draws unused" % (name, len(cols[6]), T, len(ctx.queue)))
    out["cases"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "loader_reference.npz"), **out)


# ------------------------------------------------------------------------------------------ 2-hop construction (graph_storage.py)
class _MemColl(object):
    """in-memory stand-in for one Mongo collection: find({}) hands out the stored documents, insert_many keeps them"""
    def __init__(self):
        self.docs = []

    def find(self, q):
        import copy
        assert q == {}
        return copy.deepcopy(self.docs)

    def insert_many(self, docs):
        self.docs.extend(docs)


class _MemDB(dict):
    def __missing__(self, name):
        self[name] = _MemColl()
        return self[name]


def load_reference_graph_store(reference_root, ns):
    """exec the module-level constants and class GraphStore of code/graph_storage.py (source untouched; the module
    itself cannot be imported: it imports pymongo and matplotlib)"""
    path = os.path.join(reference_root, "code", "graph_storage.py")
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.Assign) or (isinstance(node, ast.ClassDef) and node.name == "GraphStore"):
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns["GraphStore"]


HOP2_CASES = [
    # name, n_user, n_item, slices, start, max_1hop, max_2hop, edges, users / items per collection
    ("caps_10_100", 40, 30, 5, 0, 10, 100, 2600, 20, 15),     # the reference's constants: the max_2hop cap can never bind
    ("caps_4_9_start1", 30, 24, 4, 1, 4, 9, 900, 10, 8),      # both caps bind, start_time > 0
    ("caps_3_5", 24, 20, 3, 0, 3, 5, 500, 24, 20),
]


def make_hop2(reference_root):
    from oracle import graph_ref as G
    out = {"source": np.array("reference: GraphStore.construct_coll_2hop of code/graph_storage.py:127-246 executed unmodified "
                              "(ast-extracted class; MongoDB replaced by in-memory collections; random.shuffle / "
                              "np.random.choice apply the Philox permutations of csrc/hop2.cu, call k of the reference's "
                              "own call order)")}
    names = []
    for ci, (name, nu, ni, S, start, m1, m2, edges, upc, ipc) in enumerate(HOP2_CASES):
        rng = np.random.default_rng(900 + ci)
        u1, i1 = G.random_1hop(rng, nu, ni, S, edges)
        seed = 4242 + ci
        calls = {"shuffle": 0, "choice": 0}

        class _Random(object):
            @staticmethod
            def shuffle(lst):
                perm = G.permutation(seed, G.STREAM_SHUFFLE, calls["shuffle"], len(lst))
                calls["shuffle"] += 1
                lst[:] = [lst[int(p)] for p in perm]

            @staticmethod
            def seed(x):
                pass

        class _NPRandom(object):
            @staticmethod
            def choice(a, size=None, replace=True, p=None):
                assert replace is False and p is None and size == len(a)
                perm = G.permutation(seed, G.STREAM_CHOICE, calls["choice"], len(a))
                calls["choice"] += 1
                return np.asarray(a)[perm]

        class _NP(object):
            random = _NPRandom()

            def __getattr__(self, k):
                return getattr(np, k)

        db1, db2 = _MemDB(), _MemDB()
        for i in range(nu // upc):
            db1["user_%d" % i].insert_many([{"uid": u, "1hop": [list(x) for x in u1[u]]} for u in range(i * upc + 1, (i + 1) * upc + 1)])
        for i in range(ni // ipc):
            db1["item_%d" % i].insert_many([{"iid": it, "1hop": [list(x) for x in i1[it]]}
                                            for it in range(nu + i * ipc + 1, nu + (i + 1) * ipc + 1)])
        client = {"h1": db1, "h2": db2}
        pym = type("PM", (), {"MongoClient": staticmethod(lambda url: client)})
        ns = {"pymongo": pym, "np": _NP(), "random": _Random(), "print": lambda *a, **k: None}
        GS = load_reference_graph_store(reference_root, ns)
        gs = GS(os.devnull, user_per_collection=upc, item_per_collection=ipc, start_time=start, max_1hop=m1, max_2hop=m2,
                user_num=nu, item_num=ni, db_1hop="h1", db_2hop="h2", time_slice_num=S)
        gs.construct_coll_2hop()
        user_docs = {d["uid"]: d for k, c in db2.items() if k.startswith("user_") for d in c.docs}
        item_docs = {d["iid"]: d for k, c in db2.items() if k.startswith("item_") for d in c.docs}
        assert len(user_docs) == nu and len(item_docs) == ni
        from score_b200.graph import docs_to_csr
        in_u = {u: {"1hop": u1[u], "2hop": [[] for _ in range(S)], "degrees": [[] for _ in range(S)]} for u in u1}
        in_i = {i: {"1hop": i1[i], "2hop": [[] for _ in range(S)], "degrees": [[] for _ in range(S)]} for i in i1}
        off1, ids1_in, _, _, _ = docs_to_csr(in_u, in_i, nu, ni, S)
        off1b, ids1_out, off2, ids2, deg2 = docs_to_csr(user_docs, item_docs, nu, ni, S)
        assert np.array_equal(off1, off1b)
        pre = name + "/"
        out[pre + "params"] = np.array([nu, ni, S, start, m1, m2, seed], np.int64)
        out[pre + "hop1_off"], out[pre + "hop1_ids"] = off1, ids1_in
        out[pre + "hop1_ids_out"], out[pre + "hop2_off"], out[pre + "hop2_ids"], out[pre + "hop2_deg"] = ids1_out, off2, ids2, deg2
        names.append(name)
        print("hop2 case %s: %d shuffle calls, %d choice calls, %d 2-hop ids, %d of %d 1-hop ids moved" % (
            name, calls["shuffle"], calls["choice"], len(ids2), int((ids1_in != ids1_out).sum()), len(ids1_in)))
    out["cases"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "hop2_reference.npz"), **out)


# ------------------------------------------------------------------------------------------ the reference's own graph code
WIRING_CASES = [
    # fixture name, source file, class, shape key, batch kwargs
    ("score", "code/score/score.py", "SCORE", "tiny", dict(seed=201, batch=10)),
    ("score_tb", "code/score/score.py", "SCORE", "tiny_tb", dict(seed=202, batch=12)),
    ("ria", "code/score/score.py", "RIA", "tiny", dict(seed=203, batch=8)),
    ("rca", "code/score/score.py", "RCA", "tiny", dict(seed=204, batch=8)),
    ("score_user", "code/score/score.py", "SCORE_USER", "tiny", dict(seed=205, batch=8)),
    ("score_item", "code/score/score.py", "SCORE_ITEM", "tiny", dict(seed=206, batch=8)),
    ("rrn", "code/slice_models/slice_model.py", "RRN", "tiny", dict(seed=207, batch=8)),
]


def load_reference_model_classes(reference_root, rel_path, shim):
    """exec the module-level constants and the model classes of the reference file (source untouched) with the TF stand-in
    bound to the names the file imports (`tf`, `GRUCell`, `np`)"""
    path = os.path.join(reference_root, rel_path)
    tree = ast.parse(open(path).read())
    ns = {"tf": shim, "GRUCell": shim.GRUCell, "np": np}
    for node in tree.body:
        if isinstance(node, (ast.Assign, ast.ClassDef)):
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


def make_wiring(reference_root):
    """the reference's OWN model classes (score.py SCORE / RIA / RCA / SCORE_USER / SCORE_ITEM, slice_model.py RRN), executed
    unmodified over tools/tf_shim.py: variable names / shapes / creation order, eval() through the reference's own eval(), the
    gradient of the loss node the reference built, two optimizer steps through its [loss, train_step] fetch, and one call of
    its own train() (keep_prob 0.8) with the dropout masks injected"""
    import torch
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import tf_shim as shim
    from oracle import score_ref as ref
    from score_b200.synth import SHAPES, make_batch
    for name, rel, cls, shape_key, bkw in WIRING_CASES:
        shape = SHAPES[shape_key]
        cfg = ref.ScoreConfig(*shape.ctor_args(), model_type=cls)
        params = ref.init_params(cfg, PARAM_SEED, torch.float32)
        ns = load_reference_model_classes(reference_root, rel, shim)
        model = ns[cls](*shape.ctor_args())
        spec = shim.set_variables(params)
        assert spec == [(n, tuple(s)) for n, s in ref.param_specs(cfg)], "variable names / shapes / creation order differ"
        sess = shim.Session()
        batch = [x.tolist() for x in make_batch(shape, **bkw)]
        batch2 = [x.tolist() for x in make_batch(shape, **dict(bkw, seed=bkw["seed"] + 1000))]
        preds, labels, eloss = model.eval(sess, batch, REG)                      # the reference's own eval()

        def feed(b, lr, keep):
            return {model.user_1hop_ph: b[0], model.user_2hop_ph: b[1], model.item_1hop_ph: b[2], model.item_2hop_ph: b[3],
                    model.target_user_ph: b[4], model.target_item_ph: b[5], model.label_ph: b[6], model.length_ph: b[7],
                    model.lr: lr, model.reg_lambda: REG, model.keep_prob: keep}
        loss, grads = shim.gradients(sess, model.loss, feed(batch, LR, 1.0))
        out = {"source": np.array("reference: class %s of %s executed unmodified (ast-extracted) over the TensorFlow stand-in "
                                  "tools/tf_shim.py; weights = oracle init_params(seed %d) assigned by TF variable name" % (cls, rel, PARAM_SEED)),
               "shape": np.array(shape_key), "model_type": np.array(cls), "param_seed": np.array(PARAM_SEED),
               "reg_lambda": np.array(REG), "lr": np.array(LR),
               "var_names": np.array([n for n, _ in spec]),
               "eval_preds": np.asarray(preds, np.float32), "eval_labels": np.asarray(labels, np.int32), "eval_loss": np.array(eloss, np.float32),
               "loss": np.array(float(loss), np.float32)}
        for i in range(8):
            out["batch/%d" % i] = np.asarray(batch[i]).astype(np.int32)
            out["batch2/%d" % i] = np.asarray(batch2[i]).astype(np.int32)
        rows, vals = ref.embedding_row_grads(grads["emb_mtx"])
        out["emb_rows"], out["emb_row_grads"] = rows.numpy().astype(np.int64), vals.numpy().astype(np.float32)
        for k, g in grads.items():
            if k != "emb_mtx":
                out["grad/" + k] = g.numpy().astype(np.float32)
        # two optimizer steps through the reference's fetch list [loss, train_step] (keep_prob fed as 1 so that the result
        # does not depend on a random mask)
        l0, _ = sess.run([model.loss, model.train_step], feed_dict=feed(batch, LR, 1.0))
        l1, _ = sess.run([model.loss, model.train_step], feed_dict=feed(batch2, LR, 1.0))
        out["train_losses"] = np.array([l0, l1], np.float32)
        for k in ("fc1/kernel", "fc3/bias", "bn1/gamma", "gru_user_side/gru_cell/gates/kernel"):
            out["after2/" + k] = shim.G.by_name[k].value.numpy().copy()
        touched = np.unique(np.concatenate([np.asarray(x).reshape(-1) for x in batch[:6]] + [np.asarray(x).reshape(-1) for x in batch2[:6]]))
        touched = touched[touched > 0][:256].astype(np.int64)
        out["after2/rows"] = touched
        out["after2/emb"] = shim.G.by_name["emb_mtx"].value.numpy()[touched].copy()
        # the reference's own train(): keep_prob 0.8 (score.py:113), masks injected so that the oracle can follow
        gen = torch.Generator().manual_seed(4000 + bkw["seed"])
        B = len(batch[6])
        masks = [(torch.rand(B, 200, generator=gen) < 0.8).float(), (torch.rand(B, 80, generator=gen) < 0.8).float()]
        shim.G.dropout_masks = iter([m.clone() for m in masks])
        l2 = model.train(sess, batch, LR, REG)
        shim.G.dropout_masks = None
        out["train_dropout_loss"] = np.array(l2, np.float32)
        out["dropout_mask1"], out["dropout_mask2"] = masks[0].numpy().astype(np.uint8), masks[1].numpy().astype(np.uint8)
        out["after3/fc1/kernel"] = shim.G.by_name["fc1/kernel"].value.numpy().copy()
        np.savez_compressed(os.path.join(OUT, "refwiring_%s.npz" % name), **out)
        print("refwiring_%s.npz: %d variables, eval loss %.6f, train losses %.6f %.6f, dropout step %.6f" % (
            name, len(spec), float(eloss), float(l0), float(l1), float(l2)))


def make_tmall(reference_root):
    """BASELINE.json config 1: the reference's bundled Tmall sample as a derived fixture (graph CSR + feature tables +
    target lines); the raw log itself stays in the reference tree"""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import tmall_sample as ts
    from score_b200.graph import docs_to_csr, feat_table
    d = ts.build(os.path.join(reference_root, "score-data", "Tmall", "raw_data", "user_log_format1.csv"))
    off1, ids1, off2, ids2, deg2 = docs_to_csr(d["user_docs"], d["item_docs"], d["n_user"], d["n_item"], d["n_slices"])
    np.savez_compressed(
        os.path.join(OUT, "tmall_sample.npz"),
        source=np.array("derived from the reference's score-data/Tmall/raw_data/user_log_format1.csv by "
                        "tools/tmall_sample.py (restated feateng_tmall.py / graph_storage.py / gen_target.py; "
                        "age and gender synthesised: user_info_format1.csv is not shipped)"),
        dims=np.array([d["n_user"], d["n_item"], d["feature_size"], d["n_slices"]], np.int64),
        hop1_off=off1, hop1_ids=ids1, hop2_off=off2, hop2_ids=ids2, hop2_deg=deg2,
        user_feat=feat_table(d["user_feat"], 1, d["n_user"], 2),
        item_feat=feat_table(d["item_feat"], d["n_user"] + 1, d["n_item"], 3),
        target_9=d["targets"][9], target_10=d["targets"][10], target_11=d["targets"][11])
    print("tmall_sample.npz: %d users, %d items, V=%d, targets %s" % (
        d["n_user"], d["n_item"], d["feature_size"], {k: len(v) for k, v in d["targets"].items()}))


def params_digest(params):
    h = hashlib.sha256()
    for k, v in params.items():
        h.update(k.encode())
        h.update(np.ascontiguousarray(v.numpy()).tobytes())
    return h.hexdigest()


MODEL_CASES = [
    # name, shape key, model type, batch kwargs, explicit lengths
    ("tiny_score", "tiny", "SCORE", dict(seed=101), None),
    ("tinytb_score", "tiny_tb", "SCORE", dict(seed=102, batch=32), None),
    ("tiny_ragged", "tiny", "SCORE", dict(seed=103, batch=9, dummy_frac=0.4), [0, 1, 2, 3, 4, 5, 6, 6, 3]),
    ("tiny_ria", "tiny", "RIA", dict(seed=104, batch=12), None),
    ("tiny_rca", "tiny", "RCA", dict(seed=105, batch=12), None),
    ("tiny_score_user", "tiny", "SCORE_USER", dict(seed=106, batch=12), None),
    ("tiny_score_item", "tiny", "SCORE_ITEM", dict(seed=107, batch=12), None),
    ("tiny_rrn", "tiny", "RRN", dict(seed=108, batch=12), None),
]
PARAM_SEED, REG, LR = 7, 1e-4, 5e-4


def make_model_case(name, shape_key, model_type, bkw, lengths):
    import torch
    from oracle import score_ref as ref
    from score_b200.synth import SHAPES, make_batch
    shape = SHAPES[shape_key]
    cfg = ref.ScoreConfig(*shape.ctor_args(), model_type=model_type)
    params = ref.init_params(cfg, PARAM_SEED, torch.float32)
    batch = list(make_batch(shape, **bkw))
    if lengths is not None:
        batch[7] = np.array(lengths, np.int32)
    tb = ref.to_batch(batch)
    loss, y, grads, inter = ref.loss_and_grads(params, tb, cfg, REG, 1.0)
    p64 = type(params)((k, v.double()) for k, v in params.items())
    loss64, y64, _, _ = ref.loss_and_grads(p64, tb, cfg, REG, 1.0)
    rows, vals = ref.embedding_row_grads(grads["emb_mtx"])
    out = {"source": np.array("oracle/score_ref.py (CPU restatement of code/score/score.py; NOT produced by TensorFlow)"),
           "shape": np.array(shape_key), "model_type": np.array(model_type), "param_seed": np.array(PARAM_SEED),
           "reg_lambda": np.array(REG), "lr": np.array(LR), "params_sha256": np.array(params_digest(params)),
           "loss": np.array(float(loss), np.float32), "loss_fp64": np.array(float(loss64)),
           "y_pred": y.numpy().astype(np.float32), "y_pred_fp64": y64.numpy(),
           "emb_rows": rows.numpy().astype(np.int64), "emb_row_grads": vals.numpy().astype(np.float32)}
    for i, x in enumerate(batch):
        out["batch/%d" % i] = np.asarray(x).astype(np.int32)
    for k, g in grads.items():
        if k != "emb_mtx":
            out["grad/" + k] = g.numpy().astype(np.float32)
    # two optimizer steps (keep_prob 1): losses and a digest-sized summary of the updated state
    orc = ref.ScoreOracle(*shape.ctor_args(), model_type=model_type, seed=PARAM_SEED)
    b2 = list(make_batch(shape, **dict(bkw, seed=bkw["seed"] + 1000)))
    if lengths is not None:
        b2[7] = np.array(lengths, np.int32)
    l0 = orc.train(None, batch, LR, REG, keep_prob=1.0)
    l1 = orc.train(None, b2, LR, REG, keep_prob=1.0)
    for i, x in enumerate(b2):
        out["batch2/%d" % i] = np.asarray(x).astype(np.int32)
    out["train_losses"] = np.array([l0, l1], np.float32)
    touched = np.unique(np.concatenate([np.asarray(x).reshape(-1) for x in batch[:6]] + [np.asarray(x).reshape(-1) for x in b2[:6]]))
    touched = touched[touched > 0][:512]
    out["after2/rows"] = touched.astype(np.int64)
    out["after2/emb"] = orc.params["emb_mtx"].numpy()[touched]
    out["after2/emb_m"] = orc.opt.m["emb_mtx"].numpy()[touched]
    out["after2/emb_v"] = orc.opt.v["emb_mtx"].numpy()[touched]
    for k in ("fc1/kernel", "fc3/bias", "bn1/gamma"):
        out["after2/" + k] = orc.params[k].numpy()
    np.savez_compressed(os.path.join(OUT, "model_%s.npz" % name), **out)
    print("model_%s.npz: loss %.6f, %d gradient rows" % (name, float(loss), len(rows)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    if not args.only or args.only == "metrics":
        make_metrics(args.reference)
    if not args.only or args.only == "loader":
        make_loader(args.reference)
    if not args.only or args.only == "tmall":
        make_tmall(args.reference)
    if not args.only or args.only == "hop2":
        make_hop2(args.reference)
    if not args.only or args.only == "wiring":
        make_wiring(args.reference)
    for c in MODEL_CASES:
        if not args.only or args.only == c[0]:
            make_model_case(*c)


if __name__ == "__main__":
    main()
