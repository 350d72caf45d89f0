"""BASELINE.json config 1: the reference's bundled Tmall sample end to end (tools/run_tmall_sample.py) - CUDA path
vs CPU oracle on identical batches and weights.  Bars (BASELINE.json): AUC within 1e-4; losses within 1e-5 relative."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_fixture_matches_the_surveyed_sample():
    """62 users; 41 / 43 / 44 target users at pred_time 9 / 10 / 11"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "tmall_sample.npz"))
    nu, ni, V, S = (int(x) for x in g["dims"])
    assert (nu, S) == (62, 14) and V == 1 + nu + ni + (V - 1 - nu - ni)
    assert [len(g["target_%d" % p]) for p in (9, 10, 11)] == [41, 43, 44]
    assert g["target_9"].shape[1] == 101
    assert g["hop1_ids"].size == 2 * 9999          # every interaction appears in one user list and one item list
    assert g["hop2_ids"].max() <= nu + ni and g["user_feat"].max() < V and g["item_feat"].max() < V


@pytest.mark.gpu
def test_tmall_sample_end_to_end_auc_parity():
    import run_tmall_sample as rt
    log = rt.run(epochs=3, with_oracle=True, verbose=False)
    lc, lo = np.asarray(log["train_loss_cuda"]), np.asarray(log["train_loss_oracle"])
    assert len(lc) == 3 and np.abs(lc - lo).max() <= 1e-5 * np.abs(lo).max()
    for split in ("validation", "test"):
        c, o = log[split]["cuda"], log[split]["oracle"]
        assert abs(c[1] - o[1]) <= 1e-4, (split, "auc", c[1], o[1])          # AUC
        assert abs(c[0] - o[0]) <= 1e-5 * abs(o[0])                          # log-loss
        assert abs(c[8] - o[8]) <= 1e-5 * abs(o[8])                          # mean batch loss incl. L2
        assert log[split]["max_pred_diff"] <= 1e-5
        assert 0.0 <= c[1] <= 1.0 and all(0.0 <= x <= 1.0 for x in c[2:8])


@pytest.mark.skipif(not os.path.isdir("/root/reference/code/score"), reason="needs the reference checkout (build container only)")
def test_config1_through_the_references_own_class_matches_the_cuda_run():
    """BASELINE.json config 1 with the reference's OWN code on the CPU - class SCORE of score.py over the TF stand-in, its
    own eval(), its own get_ranking_quality, sklearn - against the numbers the CUDA path produced on B200 for the same
    flow (profiles/r01_tmall_sample_config1.json): AUC within 1e-4 (BASELINE.json), losses / metrics within 1e-5."""
    import json
    import run_tmall_sample_reference as rr
    log = rr.run(epochs=5, verbose=False)
    cuda = json.load(open(os.path.join(ROOT, "profiles", "r01_tmall_sample_config1.json")))
    a, b = np.asarray(log["train_loss_reference"]), np.asarray(cuda["train_loss_cuda"])
    assert a.shape == b.shape and np.abs(a - b).max() <= 1e-5 * np.abs(b).max()
    for split in ("validation", "test"):
        r, c = np.asarray(log[split]["reference"]), np.asarray(cuda[split]["cuda"])
        assert abs(r[1] - c[1]) <= 1e-4, (split, "auc", r[1], c[1])
        assert np.abs(r - c).max() <= 1e-5, (split, r, c)
