#!/usr/bin/env python
"""BASELINE.json config 1 through the REFERENCE'S OWN model class, on the CPU: class SCORE of code/score/score.py executed
unmodified over the TensorFlow stand-in (tools/tf_shim.py) on the bundled Tmall sample, same flow, constants, initial
weights and batches as tools/run_tmall_sample.py (which runs the CUDA path on the GPU): 5 training steps, then the
validation / test scoring through the reference's own eval() and get_ranking_quality (train_score.py:104-163, lifted out
with ast) + sklearn.  Batches: the CPU restatement of the loader (oracle/loader_ref.py, bit-exact to the CUDA sampler and to
the reference's GraphHandler - tests/test_loader.py) with the sampler's Philox uniforms.  Dropout off on both sides, as there.

  python tools/run_tmall_sample_reference.py [--reference /root/reference] [--epochs 5]   -> JSON like run_tmall_sample.py
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

EB_DIM, HIDDEN, K, T, UF, IF = 16, 32, 10, 11, 3, 4       # train_score.py:16-17, 46-54
TRAIN_BATCH, EVAL_BATCH, LR, REG = 100, 100, 5e-4, 1e-4  # train_score.py:13-14, 371-372


def csr_to_docs(g, nu, ni, S):
    def lists(off, ids, node):
        return [ids[off[node * S + t]:off[node * S + t + 1]].tolist() for t in range(S)]
    ud = {u: {"1hop": lists(g["hop1_off"], g["hop1_ids"], u), "2hop": lists(g["hop2_off"], g["hop2_ids"], u),
              "degrees": lists(g["hop2_off"], g["hop2_deg"], u)} for u in range(1, nu + 1)}
    idocs = {i: {"1hop": lists(g["hop1_off"], g["hop1_ids"], i), "2hop": lists(g["hop2_off"], g["hop2_ids"], i),
                 "degrees": lists(g["hop2_off"], g["hop2_deg"], i)} for i in range(nu + 1, nu + ni + 1)}
    ufd = {str(u): g["user_feat"][u].tolist() for u in range(1, nu + 1)}
    ifd = {str(i): g["item_feat"][i - nu].tolist() for i in range(nu + 1, nu + ni + 1)}
    return ud, idocs, ufd, ifd


def batches(g, docs, targets, n_items, pred_time, neg, batch_size, seed, S, nu, ni):
    """the batches DeviceGraphLoader(store, batch_size, lines, 0, pred_time, neg, T, K, seed=seed) yields (score_b200/graph.py)"""
    from oracle import loader_ref as L
    ud, idocs, ufd, ifd = docs
    grp = 1 + neg
    per = batch_size // grp
    uids = [int(r[0]) for r in targets]
    iids = [[int(x) for x in r[1:1 + grp]] for r in targets]
    for bi in range((len(uids) + per - 1) // per):
        u = uids[bi * per:(bi + 1) * per]
        it = [x for row in iids[bi * per:(bi + 1) * per] for x in row]
        # time_slice_num of the LOADER is train_score.py's TIME_SLICE_NUM_Tmall = 12 (T = 12 - 0 - 1), not the 14 slices stored
        h = L.GraphHandlerRef(T + 1, ud, idocs, K, nu, ni, 0, "rs", ufd, ifd, UF, IF,
                              lambda side, ent, ts, bi=bi: L.draw_uniforms(seed, bi, side, ent, ts, T, K))
        yield L.assemble_batch(h, u, it, pred_time, 0, neg)


def run(reference_root="/root/reference", epochs=5, seed=1111, verbose=True):
    import torch
    import tf_shim as shim
    import make_golden as mg
    from oracle import score_ref as ref
    from sklearn.metrics import log_loss, roc_auc_score
    g = np.load(os.path.join(ROOT, "tests", "golden", "tmall_sample.npz"))
    nu, ni, V, S = (int(x) for x in g["dims"])
    docs = csr_to_docs(g, nu, ni, S)
    ns = mg.load_reference_model_classes(reference_root, "code/score/score.py", shim)
    model = ns["SCORE"](V, EB_DIM, HIDDEN, T, K, UF, IF)
    cfg = ref.ScoreConfig(V, EB_DIM, HIDDEN, T, K, UF, IF)
    shim.set_variables(ref.init_params(cfg, seed))
    sess = shim.Session()
    rq = mg.load_reference_metric_functions(reference_root)["get_ranking_quality"]
    log = {"train_loss_reference": []}

    def feed(b, keep):
        return {model.user_1hop_ph: b[0], model.user_2hop_ph: b[1], model.item_1hop_ph: b[2], model.item_2hop_ph: b[3],
                model.target_user_ph: b[4], model.target_item_ph: b[5], model.label_ph: b[6], model.length_ph: b[7],
                model.lr: LR, model.reg_lambda: REG, model.keep_prob: keep}
    for ep in range(epochs):
        for b in batches(g, docs, g["target_9"], 2, 9, 1, TRAIN_BATCH, seed + ep, S, nu, ni):
            loss, _ = sess.run([model.loss, model.train_step], feed_dict=feed(b, 1.0))     # train()'s fetch list, dropout off
            log["train_loss_reference"].append(float(loss))

    def evaluate(pred_time, targets):
        preds, labels, iids, losses = [], [], [], []
        for b in batches(g, docs, targets, 100, pred_time, 99, EVAL_BATCH, seed + 1000, S, nu, ni):
            p, lab, loss = model.eval(sess, b, REG)                                        # the reference's own eval()
            preds += p; labels += lab; losses.append(float(loss))
            iids += np.asarray(b[5])[:, 0].tolist()                                        # train_score.py:157
        q = rq(preds, iids)                                                                # train_score.py:122-142
        return [float(log_loss(labels, preds)), float(roc_auc_score(labels, preds))] + [float(x) for x in q] + [sum(losses) / len(losses)]
    log["validation"] = {"reference": evaluate(10, g["target_10"])}
    log["test"] = {"reference": evaluate(11, g["target_11"])}
    log["names"] = ["logloss", "auc", "ndcg5", "ndcg10", "hr1", "hr5", "hr10", "mrr", "loss"]
    if verbose:
        print(json.dumps(log, indent=1))
    return log


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--epochs", type=int, default=5)
    a = ap.parse_args()
    run(a.reference, a.epochs)
