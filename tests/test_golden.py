"""Checks against the committed fixtures of tests/golden/ (made by tools/make_golden.py).

* metrics_reference.npz holds outputs of the reference's OWN functions (code/score/train_score.py:104-142, executed
  unmodified in the build container) - it pins oracle/metrics_ref.py (CPU) and the CUDA metrics kernel (GPU).
* model_*.npz hold outputs of the CPU restatement (TensorFlow is not installable: parity unpinned); they freeze the
  oracle and let the GPU box check the CUDA path against stored vectors.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import metrics_ref
from oracle import score_ref as ref
from score_b200.synth import SHAPES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CUDA_MODEL_TYPES = ("SCORE", "RIA", "RCA", "SCORE_USER", "SCORE_ITEM", "RRN")


def _load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def _model_cases():
    return sorted(os.path.basename(p)[len("model_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN, "model_*.npz")))


def _batch(g, prefix="batch"):
    return tuple(g["%s/%d" % (prefix, i)] for i in range(8))


def _params(g):
    shape = SHAPES[str(g["shape"])]
    cfg = ref.ScoreConfig(*shape.ctor_args(), model_type=str(g["model_type"]))
    return shape, cfg, ref.init_params(cfg, int(g["param_seed"]), torch.float32)


# ------------------------------------------------------------------------------------------ CPU: oracle vs fixtures
def test_fixture_inventory():
    assert os.path.exists(os.path.join(GOLDEN, "metrics_reference.npz"))
    assert len(_model_cases()) >= 8
    assert str(_load("metrics_reference.npz")["source"]).startswith("reference:")


def test_metrics_oracle_matches_the_references_own_functions():
    g = _load("metrics_reference.npz")
    for case in g["cases"]:
        want = g[case + "/expect"]
        got = metrics_ref.eval_metrics(g[case + "/preds"], g[case + "/labels"], g[case + "/iids"])
        np.testing.assert_allclose(np.array(got), want, rtol=1e-13, atol=0, err_msg=str(case))
    rl = list(range(10, 30))
    ts = (10, 12, 14, 15, 99)
    assert [metrics_ref.getNDCG_at_K(rl, t, k) for t in ts for k in (5, 10)] == g["helpers/ndcg"].tolist()
    assert [metrics_ref.getHR_at_K(rl, t, k) for t in ts for k in (1, 5, 10)] == g["helpers/hr"].tolist()
    assert [metrics_ref.getMRR(rl, t) for t in (10, 12, 14, 15, 29, 99)] == g["helpers/mrr"].tolist()


@pytest.mark.parametrize("case", ["tiny_score", "tiny_ragged", "tiny_rca", "tiny_rrn"])
def test_model_oracle_reproduces_its_fixture(case):
    """The restatement has not drifted since the fixture was written (same seeds -> same numbers)."""
    g = _load("model_%s.npz" % case)
    shape, cfg, params = _params(g)
    import hashlib
    h = hashlib.sha256()
    for k, v in params.items():
        h.update(k.encode())
        h.update(np.ascontiguousarray(v.numpy()).tobytes())
    assert h.hexdigest() == str(g["params_sha256"])
    loss, y, grads, _ = ref.loss_and_grads(params, ref.to_batch(_batch(g)), cfg, float(g["reg_lambda"]), 1.0)
    assert float(loss) == pytest.approx(float(g["loss"]), rel=2e-6)
    assert abs(float(g["loss"]) - float(g["loss_fp64"])) <= 1e-5 * abs(float(g["loss_fp64"]))
    np.testing.assert_allclose(y.numpy(), g["y_pred"], rtol=1e-5, atol=1e-7)
    rows, vals = ref.embedding_row_grads(grads["emb_mtx"])
    assert np.array_equal(rows.numpy(), g["emb_rows"])
    scale = np.abs(g["emb_row_grads"]).max()
    assert np.abs(vals.numpy() - g["emb_row_grads"]).max() <= 1e-5 * scale
    for k in grads:
        if k == "emb_mtx":
            continue
        want = g["grad/" + k]
        assert np.abs(grads[k].numpy() - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-6), k


# ------------------------------------------------------------------------------------------ GPU: CUDA path vs fixtures
@pytest.mark.gpu
def test_cuda_metrics_match_the_references_own_functions():
    from score_b200 import model as sb
    g = _load("metrics_reference.npz")
    m = sb.SCORE(*SHAPES["tiny"].ctor_args(), use_graph=False)
    for case in g["cases"]:
        want = g[case + "/expect"]
        got = np.array(m.eval_metrics(g[case + "/preds"], g[case + "/iids"], g[case + "/labels"], 100))
        assert got[0] == pytest.approx(want[0], rel=1e-12), case      # logloss
        assert got[1] == pytest.approx(want[1], abs=1e-12), case      # AUC (bar 1e-4)
        # in every fixture the positive's score is untied within its group, so its rank does not depend on the tie rule
        np.testing.assert_allclose(got[2:8], want[2:8], rtol=1e-12, atol=1e-15, err_msg=str(case))
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", _model_cases())
def test_cuda_forward_backward_matches_fixture(case):
    from score_b200 import model as sb
    g = _load("model_%s.npz" % case)
    mt = str(g["model_type"])
    if mt not in CUDA_MODEL_TYPES:
        pytest.skip("model type %s not on the CUDA path" % mt)
    shape, cfg, params = _params(g)
    m = getattr(sb, mt)(*shape.ctor_args(), adam_mode="dense", init_weights=False, use_graph=False, seed=7)
    m.load_params(params)
    loss = m.forward_backward(_batch(g), float(g["reg_lambda"]), 1.0)
    assert loss == pytest.approx(float(g["loss_fp64"]), rel=1e-5)
    y = m.get_buffer("y_pred")
    assert np.abs(y - g["y_pred_fp64"]).max() <= 1e-5 * np.abs(g["y_pred_fp64"]).max()
    rows, vals = m.embedding_row_grads()
    assert np.array_equal(rows, g["emb_rows"])                       # gradient row set: bit-exact
    assert np.abs(vals - g["emb_row_grads"]).max() <= 1e-5 * np.abs(g["emb_row_grads"]).max()
    for name, _ in m.tensor_names():
        if name == "emb_mtx" or name in ref.NON_TRAINABLE:
            continue
        want = g["grad/" + name].reshape(-1)
        scale = np.abs(want).max()
        if name.endswith("/bias"):
            scale = max(scale, np.abs(g["grad/" + name[:-5] + "/kernel"]).max())
        got = m.get_buffer("grad/" + name)
        assert np.abs(got - want).max() <= 1e-5 * max(scale, 1e-30), name
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", _model_cases())
@pytest.mark.parametrize("adam_mode", ["dense", "lazy"])
def test_cuda_two_train_steps_match_fixture(case, adam_mode):
    from score_b200 import model as sb
    g = _load("model_%s.npz" % case)
    mt = str(g["model_type"])
    if mt not in CUDA_MODEL_TYPES:
        pytest.skip("model type %s not on the CUDA path" % mt)
    shape, cfg, params = _params(g)
    m = getattr(sb, mt)(*shape.ctor_args(), adam_mode=adam_mode, init_weights=False, use_graph=False, seed=7)
    m.load_params(params)
    l0 = m.train(None, _batch(g), float(g["lr"]), float(g["reg_lambda"]), keep_prob=1.0)
    l1 = m.train(None, _batch(g, "batch2"), float(g["lr"]), float(g["reg_lambda"]), keep_prob=1.0)
    np.testing.assert_allclose([l0, l1], g["train_losses"], rtol=1e-5)
    rows = g["after2/rows"]
    emb = m.get_tensor("emb_mtx")[rows]
    # two Adam steps move a weight by ~2*lr whatever the gradient's size, so compare at the scale of that movement
    assert np.abs(emb - g["after2/emb"]).max() <= 2e-2 * float(g["lr"])
    ev = m.get_tensor("emb_mtx/Adam_1")[rows]
    assert np.abs(ev - g["after2/emb_v"]).max() <= 2e-5 * max(np.abs(g["after2/emb_v"]).max(), 1e-30)
    em = m.get_tensor("emb_mtx/Adam")[rows]
    assert np.abs(em - g["after2/emb_m"]).max() <= 2e-5 * max(np.abs(g["after2/emb_m"]).max(), 1e-30)
    for k in ("fc1/kernel", "fc3/bias", "bn1/gamma"):
        assert np.abs(m.get_tensor(k).reshape(-1) - g["after2/" + k].reshape(-1)).max() <= 2e-2 * float(g["lr"]), k
    m.close()
