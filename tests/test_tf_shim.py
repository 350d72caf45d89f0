"""The TF-1.x stand-in (tools/tf_shim.py) against the constants TensorFlow's own unit tests publish.

The reference-wiring fixtures (tests/golden/refwiring_*.npz, tests/test_reference_wiring.py) are produced by running the
reference's model classes over this stand-in, so the stand-in's op semantics are part of the parity chain.  The same
published cases that pin the oracle (tests/test_tf_known_answers.py) are run here through the stand-in's GRAPH API -
placeholder -> op -> Session.run, AdamOptimizer.minimize - the way score.py drives TensorFlow."""
import math
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import tf_shim as tf   # noqa: E402

from test_tf_known_answers import (TF_GRU_CASES, TF_LOGLOSS_LABEL, TF_LOGLOSS_PRED, adam_update_numpy, np_batch_norm,   # noqa: E402
                                   tf_logloss_expected)


@pytest.fixture(autouse=True)
def fresh_graph():
    tf.reset_default_graph()
    tf.G.dtype = torch.float64
    yield
    tf.reset_default_graph()


def assign(values):
    for v in tf.global_variables():
        n = v.name[:-2]
        if n in values:
            v.value = torch.as_tensor(np.asarray(values[n], np.float64)).reshape(v.value.shape).clone()


@pytest.mark.parametrize("x,h,want", TF_GRU_CASES)
def test_dynamic_rnn_gru_cell_reproduces_rnn_cell_test(x, h, want):
    """rnn_cell_test.testGRUCell (0.175991 / 0.156736) as the second step of tf.nn.dynamic_rnn(GRUCell(2)): the first
    step is crafted to leave h = 0.1 in the state (see tests/test_tf_known_answers.py)."""
    D, H = x.shape[1], h.shape[1]
    Dx = D + 1
    inp = tf.placeholder(tf.float32, [None, 2, Dx])
    length = tf.placeholder(tf.int32, [None, ])
    out, last = tf.nn.dynamic_rnn(tf.GRUCell(H), inputs=inp, sequence_length=length, dtype=tf.float32, scope="gru_t")
    names = [v.name for v in tf.global_variables()]
    assert names == ["gru_t/gru_cell/gates/kernel:0", "gru_t/gru_cell/gates/bias:0",
                     "gru_t/gru_cell/candidate/kernel:0", "gru_t/gru_cell/candidate/bias:0"]    # TF's names and order
    gk = np.zeros((Dx + H, 2 * H)); ck = np.zeros((Dx + H, H))
    gk[:D] = 0.5; gk[Dx:] = 0.5; ck[:D] = 0.5; ck[Dx:] = 0.5
    gk[D, H:] = -1.0
    ck[D, :] = math.atanh(0.2)
    assign({"gru_t/gru_cell/gates/kernel": gk, "gru_t/gru_cell/gates/bias": np.ones(2 * H),
            "gru_t/gru_cell/candidate/kernel": ck, "gru_t/gru_cell/candidate/bias": np.zeros(H)})
    xs = np.zeros((1, 2, Dx)); xs[0, 0, D] = 1.0; xs[0, 1, :D] = x[0]
    o, l = tf.Session().run([out, last], {inp: xs, length: [2]})
    assert np.allclose(o[0, 0], 0.1, atol=1e-12)
    assert np.allclose(o[0, 1], want, rtol=1e-6, atol=1e-6) and np.allclose(l[0], want, rtol=1e-6, atol=1e-6)
    # beyond sequence_length: outputs zero, state copied through (dynamic_rnn's documented behaviour)
    o1, l1 = tf.Session().run([out, last], {inp: xs, length: [1]})
    assert np.allclose(o1[0, 1], 0.0) and np.allclose(l1[0], 0.1, atol=1e-12)
    assert out.get_shape().as_list() == [None, 2, H]


def test_adam_optimizer_minimize_reproduces_adam_test_basic():
    """adam_test.testBasic: var0 = [1, 2], var1 = [3, 4], constant gradients 0.1 / 0.01, three steps of
    AdamOptimizer(0.001) - here through minimize() of a loss whose gradient is that constant, fetched the way
    score.py:105 does (sess.run([loss, train_step]))."""
    v0 = tf.get_variable("var0", [2]); v1 = tf.get_variable("var1", [2])
    assign({"var0": [1.0, 2.0], "var1": [3.0, 4.0]})
    lr = tf.placeholder(tf.float32, [])
    loss = tf.reduce_sum(v0 * 0.1) + tf.reduce_sum(v1 * 0.01)
    step = tf.train.AdamOptimizer(learning_rate=lr).minimize(loss)
    sess = tf.Session()
    p0, p1 = np.array([1.0, 2.0]), np.array([3.0, 4.0])
    m0 = v0n = m1 = v1n = 0.0
    for t in range(1, 4):
        pre, _ = sess.run([loss, step], {lr: 0.001})
        assert pre == pytest.approx(0.1 * p0.sum() + 0.01 * p1.sum(), rel=1e-12)      # the loss of the PRE-update variables
        p0, m0, v0n = adam_update_numpy(p0, np.array([0.1, 0.1]), t, m0, v0n)
        p1, m1, v1n = adam_update_numpy(p1, np.array([0.01, 0.01]), t, m1, v1n)
        assert np.allclose(v0.value.numpy(), p0, rtol=1e-12) and np.allclose(v1.value.numpy(), p1, rtol=1e-12)


def test_log_loss_reproduces_losses_test():
    y = tf.placeholder(tf.float32, [None, ]); lab = tf.placeholder(tf.int32, [None, ])
    node = tf.losses.log_loss(lab, y)
    got = tf.Session().run(node, {y: TF_LOGLOSS_PRED, lab: TF_LOGLOSS_LABEL})
    assert float(got) == pytest.approx(tf_logloss_expected(), rel=1e-12)
    assert float(tf.Session().run(node, {y: TF_LOGLOSS_LABEL, lab: TF_LOGLOSS_LABEL})) == pytest.approx(0.0, abs=1e-3)


def test_batch_normalization_inference_reproduces_np_batch_norm():
    x = tf.placeholder(tf.float32, [None, 5])
    out = tf.layers.batch_normalization(inputs=x, name="bn1")
    second = tf.layers.batch_normalization(inputs=x)
    third = tf.layers.batch_normalization(inputs=x)
    names = [v.name for v in tf.global_variables()]
    assert names[:4] == ["bn1/gamma:0", "bn1/beta:0", "bn1/moving_mean:0", "bn1/moving_variance:0"]
    assert names[4].startswith("batch_normalization/") and names[8].startswith("batch_normalization_1/")   # TF's auto-numbering
    assert [v.name for v in tf.trainable_variables() if v.name.startswith("bn1")] == ["bn1/gamma:0", "bn1/beta:0"]
    rng = np.random.default_rng(0)
    g, b, m, v = rng.random(5) + 0.5, rng.standard_normal(5), rng.standard_normal(5) * 0.1, rng.random(5) + 0.5
    assign({"bn1/gamma": g, "bn1/beta": b, "bn1/moving_mean": m, "bn1/moving_variance": v})
    xs = rng.standard_normal((3, 5))
    got = tf.Session().run(out, {x: xs})
    assert np.allclose(got, np_batch_norm(xs, m, v, b, g, 1e-3), rtol=1e-12, atol=1e-12)
    assert second is not third


def test_l2_loss_sequence_mask_softmax_where():
    x = tf.placeholder(tf.float32, [None, 2])
    assert float(tf.Session().run(tf.nn.l2_loss(x), {x: [[1.0, 0.0], [3.0, 2.0]]})) == pytest.approx(7.0)     # nn_test.testL2Loss
    ln = tf.placeholder(tf.int32, [None, ])
    mask = tf.Session().run(tf.sequence_mask(ln, 5, dtype=tf.float32), {ln: [1, 3, 2]})                       # docstring example
    assert np.array_equal(mask, [[1, 0, 0, 0, 0], [1, 1, 1, 0, 0], [1, 1, 0, 0, 0]])
    # score.py:178-183: masked scores get -2**32 + 1 before the softmax -> exactly zero weight
    sc = tf.placeholder(tf.float32, [None, 3]); mk = tf.placeholder(tf.float32, [None, 3])
    pad = tf.ones_like(sc) * (-2 ** 32 + 1)
    w = tf.nn.softmax(tf.where(tf.equal(mk, tf.ones_like(mk)), sc, pad))
    got = tf.Session().run(w, {sc: [[1.0, 2.0, 3.0]], mk: [[1.0, 1.0, 0.0]]})
    e = np.exp([1.0, 2.0])
    assert np.allclose(got[0], [e[0] / e.sum(), e[1] / e.sum(), 0.0], rtol=1e-12)


def test_dense_layers_are_numbered_like_tf_layers():
    """tf.layers.dense without a name: 'dense', 'dense_1', ... in creation order; kernel before bias; use_bias=False
    creates no bias (the co-attention projections of score.py:150-153)."""
    x = tf.placeholder(tf.float32, [None, 4])
    a = tf.layers.dense(x, 3, activation=tf.nn.relu)
    b = tf.layers.dense(a, 2, use_bias=False)
    c = tf.layers.dense(b, 1, name="fc3")
    assert [v.name for v in tf.global_variables()] == ["dense/kernel:0", "dense/bias:0", "dense_1/kernel:0", "fc3/kernel:0", "fc3/bias:0"]
    rng = np.random.default_rng(1)
    k0, b0, k1, k3, b3 = rng.standard_normal((4, 3)), rng.standard_normal(3), rng.standard_normal((3, 2)), rng.standard_normal((2, 1)), rng.standard_normal(1)
    assign({"dense/kernel": k0, "dense/bias": b0, "dense_1/kernel": k1, "fc3/kernel": k3, "fc3/bias": b3})
    xs = rng.standard_normal((2, 4))
    want = (np.maximum(xs @ k0 + b0, 0) @ k1) @ k3 + b3
    assert np.allclose(tf.Session().run(c, {x: xs}), want, rtol=1e-12)
    assert c.get_shape().as_list() == [None, 1]


def test_dropout_keep_prob_one_is_identity_and_masks_scale():
    x = tf.placeholder(tf.float32, [None, 4]); kp = tf.placeholder(tf.float32, [])
    y = tf.nn.dropout(x, keep_prob=kp)
    xs = np.arange(8.0).reshape(2, 4)
    assert np.array_equal(tf.Session().run(y, {x: xs, kp: 1.0}), xs)
    m = torch.tensor([[1, 0, 1, 0], [0, 1, 1, 1]])
    tf.G.dropout_masks = iter([m])
    got = tf.Session().run(y, {x: xs, kp: 0.8})
    assert np.allclose(got, xs * m.numpy() / 0.8)         # tf.nn.dropout docstring: kept units scaled by 1 / keep_prob
