"""The oracle against the reference's OWN model classes.

tests/golden/refwiring_*.npz were produced by executing class SCORE / RIA / RCA / SCORE_USER / SCORE_ITEM of
code/score/score.py and class RRN of code/slice_models/slice_model.py UNMODIFIED (source lifted out with ast) over a
stand-in for the TensorFlow-1.x ops they call (tools/tf_shim.py; tools/make_golden.py: make_wiring).  The reference's
source decides the wiring - which ids are looked up, what is tiled / concatenated / fed to which layer, the variable
names, shapes and creation order, what eval() and train() feed and fetch; the stand-in supplies the op semantics, which
are pinned separately to TensorFlow's published unit-test constants (tests/test_tf_known_answers.py).  The oracle
(oracle/score_ref.py), given the same weights and batches, must reproduce those outputs; the CUDA path is compared with
the oracle on the GPU (tests/test_parity_gpu.py), so the chain reference source -> oracle -> CUDA is closed."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import score_ref as ref
from score_b200.synth import SHAPES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[len("refwiring_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN, "refwiring_*.npz")))


def _load(case):
    g = np.load(os.path.join(GOLDEN, "refwiring_%s.npz" % case), allow_pickle=False)
    shape = SHAPES[str(g["shape"])]
    cfg = ref.ScoreConfig(*shape.ctor_args(), model_type=str(g["model_type"]))
    return g, shape, cfg


def _batch(g, prefix="batch"):
    return tuple(g["%s/%d" % (prefix, i)] for i in range(8))


def test_every_model_class_of_the_path_has_a_reference_wiring_fixture():
    types = set()
    for c in CASES:
        g, _, _ = _load(c)
        assert str(g["source"]).startswith("reference:")
        types.add(str(g["model_type"]))
    assert types == set(ref.MODEL_TYPES)


@pytest.mark.parametrize("case", CASES)
def test_variable_names_and_creation_order_are_the_references(case):
    """tf.layers.dense auto-numbering ('dense', 'dense_1', ...), GRU scopes, bn1 / fc names: the list the reference's
    constructor actually creates, in its order"""
    g, shape, cfg = _load(case)
    assert [str(n) for n in g["var_names"]] == [n for n, _ in ref.param_specs(cfg)]


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_the_references_eval_and_gradients(case):
    g, shape, cfg = _load(case)
    params = ref.init_params(cfg, int(g["param_seed"]), torch.float32)
    reg = float(g["reg_lambda"])
    orc = ref.ScoreOracle(*shape.ctor_args(), model_type=cfg.model_type, seed=int(g["param_seed"]))
    preds, labels, eloss = orc.eval(None, _batch(g), reg)                  # vs the reference's own eval()
    np.testing.assert_allclose(preds, g["eval_preds"], rtol=2e-6, atol=1e-7)
    assert labels == g["eval_labels"].tolist()
    assert eloss == pytest.approx(float(g["eval_loss"]), rel=2e-6)
    loss, y, grads, _ = ref.loss_and_grads(params, ref.to_batch(_batch(g)), cfg, reg, 1.0)
    assert float(loss) == pytest.approx(float(g["loss"]), rel=2e-6)
    rows, vals = ref.embedding_row_grads(grads["emb_mtx"])
    assert np.array_equal(rows.numpy(), g["emb_rows"])                    # gradient row set
    assert np.abs(vals.numpy() - g["emb_row_grads"]).max() <= 2e-6 * np.abs(g["emb_row_grads"]).max()
    for k, gr in grads.items():
        if k == "emb_mtx":
            continue
        want = g["grad/" + k]
        scale = np.abs(want).max()
        if k.endswith("/bias"):
            scale = max(scale, np.abs(g["grad/" + k[:-5] + "/kernel"]).max())
        assert np.abs(gr.numpy() - want).max() <= 5e-6 * max(scale, 1e-30), k


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_the_references_train_steps(case):
    """two optimizer steps through the reference's [loss, train_step] fetch (keep_prob fed as 1), then one call of the
    reference's own train() - keep_prob 0.8, score.py:113 - with the dropout masks injected on both sides"""
    g, shape, cfg = _load(case)
    lr, reg = float(g["lr"]), float(g["reg_lambda"])
    orc = ref.ScoreOracle(*shape.ctor_args(), model_type=cfg.model_type, seed=int(g["param_seed"]))
    l0 = orc.train(None, _batch(g), lr, reg, keep_prob=1.0)
    l1 = orc.train(None, _batch(g, "batch2"), lr, reg, keep_prob=1.0)
    np.testing.assert_allclose([l0, l1], g["train_losses"], rtol=3e-6)
    for k in ("fc1/kernel", "fc3/bias", "bn1/gamma", "gru_user_side/gru_cell/gates/kernel"):
        assert np.abs(orc.params[k].numpy() - g["after2/" + k]).max() <= 1e-2 * lr, k     # two Adam steps move a weight by ~2 lr
    rows = g["after2/rows"]
    assert np.abs(orc.params["emb_mtx"].numpy()[rows] - g["after2/emb"]).max() <= 1e-2 * lr
    masks = (torch.from_numpy(g["dropout_mask1"].astype(np.float32)), torch.from_numpy(g["dropout_mask2"].astype(np.float32)))
    l2 = orc.train(None, _batch(g), lr, reg, keep_prob=0.8, dropout_masks=masks)
    assert l2 == pytest.approx(float(g["train_dropout_loss"]), rel=5e-6)
    assert np.abs(orc.params["fc1/kernel"].numpy() - g["after3/fc1/kernel"]).max() <= 1e-2 * lr


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_path_matches_the_references_own_classes(case):
    """the CUDA path directly against the outputs of the reference's classes.  This is the WIRING check - an error there is
    O(1) - so the bar is 1e-4; the 1e-5 bar itself is held against the oracle (tests/test_parity_gpu.py), and the oracle
    equals these fixtures to 1e-7 (tests above)"""
    from score_b200 import model as sb
    g, shape, cfg = _load(case)
    params = ref.init_params(cfg, int(g["param_seed"]), torch.float32)
    reg = float(g["reg_lambda"])
    m = getattr(sb, cfg.model_type)(*shape.ctor_args(), adam_mode="dense", init_weights=False, use_graph=False, seed=7)
    assert [n for n, _ in m.tensor_names()] == [str(n) for n in g["var_names"]]       # same variables, same order
    m.load_params(params)
    preds, labels, eloss = m.eval(None, _batch(g), reg)
    np.testing.assert_allclose(preds, g["eval_preds"], rtol=1e-4, atol=1e-6)
    assert labels == g["eval_labels"].tolist()
    assert eloss == pytest.approx(float(g["eval_loss"]), rel=1e-4)
    loss = m.forward_backward(_batch(g), reg, 1.0)
    assert loss == pytest.approx(float(g["loss"]), rel=1e-4)
    rows, vals = m.embedding_row_grads()
    assert np.array_equal(rows, g["emb_rows"])
    assert np.abs(vals - g["emb_row_grads"]).max() <= 1e-4 * np.abs(g["emb_row_grads"]).max()
    for name, _ in m.tensor_names():
        if name == "emb_mtx" or name in ref.NON_TRAINABLE:
            continue
        want = g["grad/" + name].reshape(-1)
        scale = np.abs(want).max()
        if name.endswith("/bias"):
            scale = max(scale, np.abs(g["grad/" + name[:-5] + "/kernel"]).max())
        assert np.abs(m.get_buffer("grad/" + name) - want).max() <= 1e-4 * max(scale, 1e-30), name
    lr = float(g["lr"])
    l0 = m.train(None, _batch(g), lr, reg, keep_prob=1.0)
    l1 = m.train(None, _batch(g, "batch2"), lr, reg, keep_prob=1.0)
    np.testing.assert_allclose([l0, l1], g["train_losses"], rtol=1e-4)
    m.close()


@pytest.mark.skipif(not os.path.isdir("/root/reference/code/score"), reason="needs the reference checkout (build container only)")
def test_fixtures_regenerate_from_the_reference_source(tmp_path, monkeypatch):
    """run the generator again - the reference's classes are lifted out of /root/reference and executed over the stand-in
    right now - and compare with the committed fixtures (the GPU box has no reference checkout: it uses the files)"""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_golden_mod", os.path.join(root, "tools", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    monkeypatch.setattr(mg, "OUT", str(tmp_path))
    monkeypatch.setattr(mg, "WIRING_CASES", [c for c in mg.WIRING_CASES if c[0] in ("score", "rrn")])
    mg.make_wiring("/root/reference")
    for name in ("score", "rrn"):
        new = np.load(os.path.join(str(tmp_path), "refwiring_%s.npz" % name), allow_pickle=False)
        old = np.load(os.path.join(GOLDEN, "refwiring_%s.npz" % name), allow_pickle=False)
        assert sorted(new.files) == sorted(old.files)
        for k in old.files:
            if old[k].dtype.kind in "fc":
                np.testing.assert_allclose(new[k], old[k], rtol=1e-6, atol=1e-9, err_msg=k)
            else:
                assert np.array_equal(new[k], old[k]), k


GEOMETRIES = [
    # name, model class, overrides of the tiny shape, ragged lengths
    ("ccmr_widths", "SCORE", dict(user_fnum=1, item_fnum=5, max_time_len=7, length=5), False),       # train_score.py:24-32
    ("k20", "SCORE", dict(user_fnum=1, item_fnum=5, obj_per_time_slice=20, max_time_len=4, length=3), False),
    ("wide_d64_h128", "SCORE", dict(eb_dim=64, hidden_size=128, user_fnum=1, item_fnum=1, max_time_len=5, length=4), False),
    ("tmall_t11_ragged", "SCORE", dict(max_time_len=11, length=9), True),                             # train_score.py:46-54
    ("ria_ragged", "RIA", dict(user_fnum=1, item_fnum=2, max_time_len=8, length=6), True),
    ("rca_k5", "RCA", dict(obj_per_time_slice=5), False),
    ("rrn_ccmr_widths", "RRN", dict(user_fnum=1, item_fnum=5, max_time_len=7, length=5), True),
]


@pytest.mark.skipif(not os.path.isdir("/root/reference/code/score"), reason="needs the reference checkout (build container only)")
@pytest.mark.parametrize("name,cls,over,ragged", GEOMETRIES, ids=[g[0] for g in GEOMETRIES])
def test_oracle_matches_the_references_classes_on_other_geometries(name, cls, over, ragged):
    """beyond the committed fixtures: the reference's class is built for another geometry right here (CCMR field counts,
    K = 20, d = 64 / H = 128, Tmall's T = 11, ragged lengths incl. 0 and T) over the stand-in and compared with the oracle -
    variable list, eval() through the reference's own eval(), loss and every gradient, one optimizer step"""
    import dataclasses
    import importlib.util
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import tf_shim as shim
    spec = importlib.util.spec_from_file_location("make_golden_mod2", os.path.join(root, "tools", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    from score_b200.synth import make_batch
    shape = dataclasses.replace(SHAPES["tiny"], **over)
    cfg = ref.ScoreConfig(*shape.ctor_args(), model_type=cls)
    params = ref.init_params(cfg, 77, torch.float32)
    rel = "code/slice_models/slice_model.py" if cls == "RRN" else "code/score/score.py"
    ns = mg.load_reference_model_classes("/root/reference", rel, shim)
    model = ns[cls](*shape.ctor_args())
    assert shim.set_variables(params) == [(n, tuple(s)) for n, s in ref.param_specs(cfg)]
    batch = list(make_batch(shape, seed=900 + len(name), batch=9))
    if ragged:
        batch[7] = np.array([0, shape.max_time_len, 1, 3, 2, shape.max_time_len, 4, 1, 2], np.int32)
    lists = [x.tolist() for x in batch]
    sess = shim.Session()
    reg, lr = 1e-4, 5e-4
    preds, labels, eloss = model.eval(sess, lists, reg)
    orc = ref.ScoreOracle(*shape.ctor_args(), model_type=cls, seed=77)
    o_preds, o_labels, o_loss = orc.eval(None, tuple(batch), reg)
    np.testing.assert_allclose(o_preds, preds, rtol=3e-6, atol=1e-7)
    assert o_labels == labels and o_loss == pytest.approx(eloss, rel=3e-6)
    feed = {model.user_1hop_ph: lists[0], model.user_2hop_ph: lists[1], model.item_1hop_ph: lists[2], model.item_2hop_ph: lists[3],
            model.target_user_ph: lists[4], model.target_item_ph: lists[5], model.label_ph: lists[6], model.length_ph: lists[7],
            model.lr: lr, model.reg_lambda: reg, model.keep_prob: 1.0}
    loss, grads = shim.gradients(sess, model.loss, feed)
    o_loss2, _, o_grads, _ = ref.loss_and_grads(params, ref.to_batch(tuple(batch)), cfg, reg, 1.0)
    assert float(o_loss2) == pytest.approx(float(loss), rel=3e-6)
    r_rows, r_vals = ref.embedding_row_grads(grads["emb_mtx"])
    o_rows, o_vals = ref.embedding_row_grads(o_grads["emb_mtx"])
    assert np.array_equal(o_rows.numpy(), r_rows.numpy())
    assert np.abs(o_vals.numpy() - r_vals.numpy()).max() <= 5e-6 * np.abs(r_vals.numpy()).max()
    for k, want in grads.items():
        if k == "emb_mtx":
            continue
        want = want.numpy()
        scale = np.abs(want).max()
        if k.endswith("/bias"):
            scale = max(scale, np.abs(grads[k[:-5] + "/kernel"].numpy()).max())
        assert np.abs(o_grads[k].numpy() - want).max() <= 1e-5 * max(scale, 1e-30), k
    l_ref, _ = sess.run([model.loss, model.train_step], feed_dict=feed)
    l_orc = orc.train(None, tuple(batch), lr, reg, keep_prob=1.0)
    assert l_orc == pytest.approx(float(l_ref), rel=3e-6)
    for k in ("fc1/kernel", "bn1/beta"):
        assert np.abs(orc.params[k].numpy() - shim.G.by_name[k].value.numpy()).max() <= 1e-2 * lr, k
