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
    for c in MODEL_CASES:
        if not args.only or args.only == c[0]:
            make_model_case(*c)


if __name__ == "__main__":
    main()
