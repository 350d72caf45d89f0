#!/usr/bin/env python
"""BASELINE.json config 1: SCoRe on the reference's bundled Tmall sample, end to end (train + eval AUC / ranking),
CUDA path and CPU oracle side by side on identical batches and initial weights.

The flow is train_score.py's (train() :165-258, eval() :144-163): batches of 100 from the training targets
(pred_time 9, 1 negative), evaluation on the validation / test targets (pred_time 10 / 11, 99 negatives) in batches
of 100 with sklearn-style AUC / log-loss and NDCG / HR / MRR.  Batches come from the on-GPU graph store
(DeviceGraphLoader); the oracle is fed the same ids copied to the host.  Dropout is off on both sides (TF's RNG
stream cannot be reproduced); everything else follows the reference's constants (train_score.py:46-54, 342-372).

  python tools/run_tmall_sample.py [--epochs 5] [--no-oracle]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

EB_DIM, HIDDEN, K, T, UF, IF = 16, 32, 10, 11, 3, 4       # train_score.py:16-17, 46-54
TRAIN_BATCH, EVAL_BATCH, LR, REG = 100, 100, 5e-4, 1e-4  # train_score.py:13-14, 371-372


def load_fixture():
    g = np.load(os.path.join(ROOT, "tests", "golden", "tmall_sample.npz"))
    nu, ni, V, S = (int(x) for x in g["dims"])
    return g, nu, ni, V, S


def lines_of(arr, n_items):
    return [",".join(str(int(x)) for x in row[:1 + n_items]) for row in arr]


def run(epochs=5, with_oracle=True, seed=1111, verbose=True):
    import torch
    from score_b200 import model as sb
    from score_b200.graph import DeviceGraphLoader, GraphStore
    g, nu, ni, V, S = load_fixture()
    store = GraphStore(nu, ni, S, g["hop1_off"], g["hop1_ids"], g["hop2_off"], g["hop2_ids"], g["hop2_deg"],
                       g["user_feat"], g["item_feat"], UF, IF)
    m = sb.SCORE(V, EB_DIM, HIDDEN, T, K, UF, IF, seed=seed, adam_mode="lazy", use_graph=False, init_weights=False)
    orc = None
    from oracle import score_ref as ref          # the checker; also supplies the shared initial weights
    cfg = ref.ScoreConfig(V, EB_DIM, HIDDEN, T, K, UF, IF)
    params = ref.init_params(cfg, seed)
    m.load_params(params)
    if with_oracle:
        orc = ref.ScoreOracle(V, EB_DIM, HIDDEN, T, K, UF, IF, seed=seed)
    log = {"train_loss_cuda": [], "train_loss_oracle": []}
    t0 = time.time()
    for ep in range(epochs):
        loader = DeviceGraphLoader(store, TRAIN_BATCH, lines_of(g["target_9"], 2), 0, 9, 1, T, K, seed=seed + ep)
        for batch in loader:
            host = tuple(x.cpu().numpy() for x in batch) if with_oracle else None
            log["train_loss_cuda"].append(m.train(None, batch, LR, REG, keep_prob=1.0))
            if with_oracle:
                log["train_loss_oracle"].append(orc.train(None, host, LR, REG, keep_prob=1.0))
    log["train_seconds"] = time.time() - t0

    def evaluate(pred_time, targets):
        loader = DeviceGraphLoader(store, EVAL_BATCH, lines_of(targets, 100), 0, pred_time, 99, T, K, seed=seed + 1000)
        pc, po, labels, iids, lc, lo = [], [], [], [], [], []
        for batch in loader:
            host = tuple(x.cpu().numpy() for x in batch)
            p, lab, loss = m.eval(None, batch, REG)
            pc += p; labels += lab; lc.append(loss)
            iids += host[5][:, 0].tolist()                       # train_score.py:157
            if with_oracle:
                p2, _, loss2 = orc.eval(None, host, REG)
                po += p2; lo.append(loss2)
        out = {"cuda": list(m.eval_metrics(pc, iids, labels, 100)) + [sum(lc) / len(lc)]}
        if with_oracle:
            from oracle import metrics_ref
            out["oracle"] = list(metrics_ref.eval_metrics(po, labels, iids, 100, stable=True)) + [sum(lo) / len(lo)]
            out["max_pred_diff"] = float(np.abs(np.asarray(pc) - np.asarray(po)).max())
        return out
    log["validation"] = evaluate(10, g["target_10"])
    log["test"] = evaluate(11, g["target_11"])
    log["names"] = ["logloss", "auc", "ndcg5", "ndcg10", "hr1", "hr5", "hr10", "mrr", "loss"]
    store.close(); m.close()
    if verbose:
        print(json.dumps(log, indent=1))
    return log


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--no-oracle", action="store_true")
    a = ap.parse_args()
    run(a.epochs, not a.no_oracle)
