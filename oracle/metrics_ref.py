"""CPU oracle for the evaluation arithmetic of code/score/train_score.py.  TEST INFRASTRUCTURE ONLY.

Restates, expression for expression, getNDCG_at_K / getHR_at_K / getMRR (train_score.py:104-120),
get_ranking_quality (train_score.py:122-142) and the metric lines of eval() (train_score.py:158-161:
sklearn log_loss and roc_auc_score over all predictions).  PARITY UNPINNED at the reference boundary
(the reference has no tests); the formulas are pinned by hand-derived known answers in
tests/test_oracle.py (KA-5 of SURVEY.md section 4).
"""
import math

import numpy as np
from sklearn.metrics import log_loss, roc_auc_score

TEST_NEG_SAMPLE_NUM = 99


def getNDCG_at_K(ranklist, target_item, k):
    for i in range(k):
        if ranklist[i] == target_item:
            return math.log(2) / math.log(i + 2)
    return 0


def getHR_at_K(ranklist, target_item, k):
    return 1 if target_item in ranklist[:k] else 0


def getMRR(ranklist, target_item):
    for i in range(len(ranklist)):
        if ranklist[i] == target_item:
            return 1. / (i + 1)
    return 0


def get_ranking_quality(preds, target_iids, group=TEST_NEG_SAMPLE_NUM + 1, stable=False):
    """train_score.py:122-142.  ``stable=True`` replaces np.argsort's default (unstable) sort by the stable one:
    the tie rule the CUDA kernel documents.  On tie-free groups both give the same rank list."""
    preds = np.array(preds).reshape(-1, group).tolist()
    target_iids = np.array(target_iids).reshape(-1, group).tolist()
    pos_iids = np.array(target_iids).reshape(-1, group)[:, 0].flatten().tolist()
    out = [[] for _ in range(6)]
    for i in range(len(preds)):
        order = np.argsort(preds[i], kind="stable") if stable else np.argsort(preds[i])
        ranklist = list(reversed(np.take(target_iids[i], order)))
        t = pos_iids[i]
        out[0].append(getNDCG_at_K(ranklist, t, 5))
        out[1].append(getNDCG_at_K(ranklist, t, 10))
        out[2].append(getHR_at_K(ranklist, t, 1))
        out[3].append(getHR_at_K(ranklist, t, 5))
        out[4].append(getHR_at_K(ranklist, t, 10))
        out[5].append(getMRR(ranklist, t))
    return tuple(float(np.mean(v)) for v in out)


def eval_metrics(preds, labels, target_iids, group=TEST_NEG_SAMPLE_NUM + 1, stable=False):
    """(logloss, auc, ndcg5, ndcg10, hr1, hr5, hr10, mrr) as eval() computes them (train_score.py:158-161)."""
    preds = [float(p) for p in preds]
    labels = [int(l) for l in labels]
    ll = log_loss(labels, preds)
    auc = roc_auc_score(labels, preds)
    return (float(ll), float(auc)) + get_ranking_quality(preds, target_iids, group, stable)
