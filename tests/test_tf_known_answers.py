"""Known-answer tests that pin the oracle's TF-1.x op semantics to values TensorFlow itself publishes.

TensorFlow 1.x cannot be installed in this image (no wheel for Python 3.12, no network), so the model
arithmetic of the oracle (oracle/score_ref.py) cannot be diffed against a TF run.  What CAN be checked
offline are the constants and NumPy reference implementations that TensorFlow's own unit tests assert
against - they are part of the published TF 1.x source tree and define the ops score.py calls:

  * tensorflow/python/kernel_tests/rnn_cell_test.py  RNNCellTest.testGRUCell      (score.py:205-208)
  * tensorflow/python/training/adam_test.py          adam_update_numpy, testBasic  (score.py:98)
  * tensorflow/python/kernel_tests/losses_test.py    LogLossTest                   (score.py:80)
  * tensorflow/python/ops/nn_batchnorm_test.py       _npBatchNorm                  (score.py:69)
  * tensorflow/python/ops/nn_test.py                 L2LossTest.testL2Loss         (score.py:94)
  * tf.sequence_mask / tf.nn.dropout docstring examples                            (score.py:192, 71-73)

Each test restates the published inputs and expected values literally and runs the ORACLE's own function
on them.  The CUDA path is checked against the same constants in tests/test_parity_gpu.py
(test_tf_known_answer_*), through the C ABI.
"""
import math

import numpy as np
import pytest
import torch

from oracle import score_ref as ref


# ------------------------------------------------------------------ rnn_cell_test.py :: testGRUCell
# with variable_scope("root", initializer=constant_initializer(0.5)):
#   x = zeros([1, 2]); m = zeros([1, 2]); g, _ = GRUCell(2)(x, m)
#   res = sess.run([g], {x: [[1., 1.]], m: [[0.1, 0.1]]});  assertAllClose(res[0], [[0.175991, 0.175991]])
# with variable_scope("other", initializer=constant_initializer(0.5)):   # input_size != num_units
#   x = zeros([1, 3]); ... {x: [[1., 1., 1.]], m: [[0.1, 0.1]]};        assertAllClose(res[0], [[0.156736, 0.156736]])
# GRUCell creates gates/bias with constant_initializer(1.0) and candidate/bias with zeros when bias_initializer is None;
# the scope initializer (0.5) reaches the two kernels only.
TF_GRU_CASES = [
    (np.array([[1.0, 1.0]]), np.array([[0.1, 0.1]]), 0.175991),
    (np.array([[1.0, 1.0, 1.0]]), np.array([[0.1, 0.1]]), 0.156736),
]


def gru_cell_once(x, h, gk, gb, ck, cb):
    """one GRUCell call through the oracle's dynamic_rnn: h enters as the state left by a crafted first step is not
    possible there (dynamic_rnn starts at zero), so the cell is evaluated on a length-1 sequence with the state
    substituted through the same formula lines (gru_dynamic_rnn's loop body)."""
    H = h.shape[1]
    value = torch.sigmoid(torch.matmul(torch.cat([x, h], 1), gk) + gb)
    r, u = value[:, :H], value[:, H:]
    c = torch.tanh(torch.matmul(torch.cat([x, r * h], 1), ck) + cb)
    return u * h + (1 - u) * c


@pytest.mark.parametrize("x,h,want", TF_GRU_CASES)
def test_gru_cell_matches_tf_rnn_cell_test(x, h, want):
    D, H = x.shape[1], h.shape[1]
    gk, gb = torch.full((D + H, 2 * H), 0.5, dtype=torch.float64), torch.ones(2 * H, dtype=torch.float64)
    ck, cb = torch.full((D + H, H), 0.5, dtype=torch.float64), torch.zeros(H, dtype=torch.float64)
    got = gru_cell_once(torch.from_numpy(x), torch.from_numpy(h), gk, gb, ck, cb)
    assert np.allclose(got.numpy(), want, rtol=1e-6, atol=1e-6)          # assertAllClose defaults
    # the same numbers through the oracle's dynamic_rnn itself: a first step crafted to leave h = 0.1 in the state
    # (u = sigmoid(0) = 0.5 and c = tanh(atanh(0.2)) = 0.2 -> h1 = 0.5 * 0 + 0.5 * 0.2), the published case as step 2.
    # The extra input column `trig` has weight 0 on every unit in step 2 (its input is 0 there).
    Dx = D + 1
    gk2 = torch.zeros(Dx + H, 2 * H, dtype=torch.float64)
    ck2 = torch.zeros(Dx + H, H, dtype=torch.float64)
    gk2[:D] = 0.5; gk2[Dx:] = 0.5; ck2[:D] = 0.5; ck2[Dx:] = 0.5
    gk2[D, H:] = -1.0                       # trigger column: update-gate pre-activation -1 + bias 1 = 0
    ck2[D, :] = math.atanh(0.2)
    xs = torch.zeros(1, 2, Dx, dtype=torch.float64)
    xs[0, 0, D] = 1.0
    xs[0, 1, :D] = torch.from_numpy(x[0])
    out, last = ref.gru_dynamic_rnn(xs, torch.tensor([2]), gk2, gb, ck2, cb, H)
    assert np.allclose(out[0, 0].numpy(), 0.1, atol=1e-12)
    assert np.allclose(out[0, 1].numpy(), want, rtol=1e-6, atol=1e-6)
    assert np.allclose(last[0].numpy(), want, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------ adam_test.py :: adam_update_numpy / testBasic
def adam_update_numpy(param, g_t, t, m, v, alpha=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
    """verbatim restatement of the NumPy reference in tensorflow/python/training/adam_test.py"""
    alpha_t = alpha * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    m_t = beta1 * m + (1 - beta1) * g_t
    v_t = beta2 * v + (1 - beta2) * g_t * g_t
    param_t = param - alpha_t * m_t / (np.sqrt(v_t) + epsilon)
    return param_t, m_t, v_t


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_adam_matches_tf_adam_test_basic(dtype):
    """testBasic: var0 = [1, 2], var1 = [3, 4], grads [0.1, 0.1] / [0.01, 0.01], 3 steps of AdamOptimizer();
    after step t: beta1_power = 0.9^(t+1), beta2_power = 0.999^(t+1), vars equal the NumPy reference."""
    npdt = np.float32 if dtype == torch.float32 else np.float64
    var0_np, var1_np = np.array([1.0, 2.0], npdt), np.array([3.0, 4.0], npdt)
    g0_np, g1_np = np.array([0.1, 0.1], npdt), np.array([0.01, 0.01], npdt)
    m0 = v0 = m1 = v1 = 0.0
    p = {"var0": torch.tensor([1.0, 2.0], dtype=dtype), "var1": torch.tensor([3.0, 4.0], dtype=dtype)}
    st = ref.AdamState(p)
    grads = {"var0": torch.from_numpy(g0_np.copy()), "var1": torch.from_numpy(g1_np.copy())}
    tol = 1e-6 if dtype == torch.float32 else 1e-12
    for t in range(1, 4):
        assert float(st.beta1_power) == pytest.approx(0.9 ** t, rel=1e-6)
        assert float(st.beta2_power) == pytest.approx(0.999 ** t, rel=1e-6)
        ref.adam_apply(p, grads, st, 0.001)
        var0_np, m0, v0 = adam_update_numpy(var0_np, g0_np, t, m0, v0)
        var1_np, m1, v1 = adam_update_numpy(var1_np, g1_np, t, m1, v1)
        assert np.allclose(p["var0"].numpy(), var0_np, rtol=tol, atol=tol)
        assert np.allclose(p["var1"].numpy(), var1_np, rtol=tol, atol=tol)
        assert np.allclose(st.m["var0"].numpy(), m0, rtol=1e-5) and np.allclose(st.v["var1"].numpy(), v1, rtol=1e-5)


# ------------------------------------------------------------------ losses_test.py :: LogLossTest
TF_LOGLOSS_PRED = np.asarray([.9, .2, .2, .8, .4, .6])
TF_LOGLOSS_LABEL = np.asarray([1.0, 0.0, 1.0, 1.0, 0.0, 0.0])


def tf_logloss_expected():
    """LogLossTest.setUp + testNonZeroLoss: -sum(labels*log(p+eps) + (1-labels)*log(1-p+eps)) / 6, eps = 1e-7"""
    eps = 1e-7
    e = np.multiply(TF_LOGLOSS_LABEL, np.log(TF_LOGLOSS_PRED + eps)) + \
        np.multiply(1 - TF_LOGLOSS_LABEL, np.log(1 - TF_LOGLOSS_PRED + eps))
    return -np.sum(e) / 6.0


def test_log_loss_matches_tf_losses_test():
    y = torch.from_numpy(TF_LOGLOSS_PRED)
    lab = torch.from_numpy(TF_LOGLOSS_LABEL)
    got = float(ref.total_loss({}, y, lab, 0.0))
    assert got == pytest.approx(tf_logloss_expected(), abs=1e-3)        # assertAlmostEqual(..., 3) in the TF test
    assert got == pytest.approx(tf_logloss_expected(), rel=1e-12)
    # testAllCorrectNoLossWeight: log_loss(labels, labels) == 0 to 3 places (the epsilon keeps it finite)
    assert float(ref.total_loss({}, lab, lab, 0.0)) == pytest.approx(0.0, abs=1e-3)


# ------------------------------------------------------------------ nn_batchnorm_test.py :: _npBatchNorm
def np_batch_norm(x, m, v, beta, gamma, epsilon):
    """_npBatchNorm with scale_after_normalization = shift_after_normalization = True"""
    y = (x - m) / np.sqrt(v + epsilon)
    y = y * gamma
    return y + beta


def test_batch_norm_inference_matches_tf_reference_formula():
    """tf.layers.batch_normalization(inputs=inp) with training left False (score.py:69): the moving statistics
    (0 / 1, never updated) normalise, epsilon = 1e-3 (the layer's default)."""
    from score_b200.synth import SHAPES, make_batch
    sh = SHAPES["tiny"]
    cfg = ref.ScoreConfig(*sh.ctor_args())
    p = ref.init_params(cfg, 4)
    gen = torch.Generator().manual_seed(0)
    p["bn1/gamma"] = torch.rand(p["bn1/gamma"].shape, generator=gen) + 0.5
    p["bn1/beta"] = torch.randn(p["bn1/beta"].shape, generator=gen)
    p["bn1/moving_mean"] = torch.randn(p["bn1/moving_mean"].shape, generator=gen) * 0.1
    p["bn1/moving_variance"] = torch.rand(p["bn1/moving_variance"].shape, generator=gen) + 0.5
    batch = ref.to_batch(make_batch(sh, seed=3))
    _, inter = ref.forward(p, batch, cfg, return_intermediates=True)
    bn = np_batch_norm(inter["fc_in"].double().numpy(), p["bn1/moving_mean"].double().numpy(),
                       p["bn1/moving_variance"].double().numpy(), p["bn1/beta"].double().numpy(),
                       p["bn1/gamma"].double().numpy(), 1e-3)
    fc1 = np.maximum(bn @ p["fc1/kernel"].double().numpy() + p["fc1/bias"].double().numpy(), 0)
    fc2 = np.maximum(fc1 @ p["fc2/kernel"].double().numpy() + p["fc2/bias"].double().numpy(), 0)
    logit = (fc2 @ p["fc3/kernel"].double().numpy() + p["fc3/bias"].double().numpy()).reshape(-1)
    assert np.allclose(logit, inter["logit"].numpy(), rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------ nn_test.py :: L2LossTest.testL2Loss
def test_l2_loss_matches_tf_nn_test():
    """x = [1, 0, 3, 2] (shape [2,2]) -> l2_loss = 7.0; build_l2norm multiplies it by reg_lambda (score.py:91-94)"""
    x = torch.tensor([[1.0, 0.0], [3.0, 2.0]])
    y = torch.tensor([0.5])
    base = float(ref.total_loss({}, y, torch.tensor([1.0]), 0.0))
    assert float(ref.total_loss({"w/kernel": x}, y, torch.tensor([1.0]), 1.0)) - base == pytest.approx(7.0, rel=1e-6)
    # the name filter of score.py:92-94: 'bias' and 'emb' variables are skipped
    assert float(ref.total_loss({"w/bias": x, "emb_mtx": x}, y, torch.tensor([1.0]), 1.0)) == pytest.approx(base)


# ------------------------------------------------------------------ docstring examples
def test_sequence_mask_docstring_example():
    """tf.sequence_mask([1, 3, 2], 5) -> [[T,F,F,F,F],[T,T,T,F,F],[T,T,F,F,F]]: the attention mask (score.py:192)
    and the dynamic_rnn length masking follow it."""
    B, T, H, D = 3, 5, 2, 2
    length = torch.tensor([1, 3, 2])
    x = torch.ones(B, T, D)
    gk, gb = torch.full((D + H, 2 * H), 0.5), torch.ones(2 * H)
    ck, cb = torch.full((D + H, H), 0.5), torch.zeros(H)
    out, _ = ref.gru_dynamic_rnn(x, length, gk, gb, ck, cb, H)
    want = np.array([[1, 0, 0, 0, 0], [1, 1, 1, 0, 0], [1, 1, 0, 0, 0]], bool)
    assert np.array_equal((out.abs().sum(-1) > 0).numpy(), want)


def test_glorot_uniform_and_truncated_normal_bounds():
    """tf.layers.dense default kernel_initializer glorot_uniform: U(-l, l), l = sqrt(6 / (fan_in + fan_out));
    tf.truncated_normal_initializer: values beyond 2 sigma are re-drawn (score.py:44)."""
    from score_b200.synth import SHAPES
    cfg = ref.ScoreConfig(*SHAPES["tiny"].ctor_args())
    p = ref.init_params(cfg, 9)
    k = p["fc1/kernel"]
    lim = math.sqrt(6.0 / (k.shape[0] + k.shape[1]))
    assert float(k.abs().max()) <= lim and float(k.abs().max()) > 0.95 * lim
    assert float(k.var()) == pytest.approx(lim * lim / 3, rel=0.1)
    e = p["emb_mtx"]
    assert float(e.abs().max()) <= 2.0 and float(e.std()) == pytest.approx(0.8796, rel=0.03)   # std of N(0,1) cut at 2
    assert float(p["gru_user_side/gru_cell/gates/bias"].min()) == 1.0
