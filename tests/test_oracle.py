"""CPU tests that pin the oracle: the reference has no golden vectors, so the
restatement is checked against hand-derivable known answers KA-1..KA-7, an fp64 twin and finite
differences.  These run everywhere (no GPU)."""
import math

import numpy as np
import pytest
import torch

from oracle import metrics_ref
from oracle import score_ref as ref
from score_b200.synth import SHAPES, make_batch

torch.manual_seed(1111)


def _tiny(model_type="SCORE"):
    sh = SHAPES["tiny"]
    cfg = ref.ScoreConfig(*sh.ctor_args(), model_type=model_type)
    return sh, cfg


def test_param_names_follow_tf_creation_order():
    _, cfg = _tiny()
    names = [n for n, _ in ref.param_specs(cfg)]
    assert names[:5] == ["emb_mtx", "dense/kernel", "dense/bias", "dense_1/kernel", "dense_1/bias"]
    assert names[5] == "gru_user_side/gru_cell/gates/kernel"
    assert "dense_5/kernel" in names and "dense_6/kernel" not in names
    assert names[-6:] == ["fc1/kernel", "fc1/bias", "fc2/kernel", "fc2/bias", "fc3/kernel", "fc3/bias"]
    # Tmall dense parameter count of SURVEY.md section 8 (128 484) pins every layer width
    tm = ref.ScoreConfig(*SHAPES["tmall"].ctor_args())
    n = sum(int(np.prod(s)) for k, s in ref.param_specs(tm) if k != "emb_mtx" and k not in ref.NON_TRAINABLE)
    assert n == 128484
    tb = ref.ScoreConfig(*SHAPES["taobao"].ctor_args())
    n = sum(int(np.prod(s)) for k, s in ref.param_specs(tb) if k != "emb_mtx" and k not in ref.NON_TRAINABLE)
    assert n == 96420


def test_ka1_co_attention_collapses_to_rank1():
    """score.py:152-153 tile both sequences on axis 3 -> seq2 weights are 1/K, seq1 weights softmax_i(r_i)."""
    B, T, K, D = 3, 4, 10, 32
    s1, s2, tg = torch.randn(B, T, K, D), torch.randn(B, T, K, D), torch.randn(B, T, D)
    W, b = torch.randn(3 * D, 1) * 0.2, torch.randn(1)
    o1, o2, info = ref.co_attention(s1, s2, tg, W, b)
    z = (tg @ W[:D]).unsqueeze(2) + s1 @ W[D:2 * D] + s2 @ W[2 * D:] + b
    r = torch.relu(z).squeeze(-1)
    w = torch.softmax(r, -1)
    assert torch.allclose(o1, (s1 * w.unsqueeze(-1)).sum(2), atol=1e-5)
    assert torch.allclose(o2, s2.mean(2), atol=1e-5)
    assert torch.allclose(info[..., :K], K * r, atol=1e-4)
    assert torch.allclose(info[..., K:], r.sum(-1, keepdim=True).expand(B, T, K), atol=1e-4)


def test_ka2_zero_id_is_zero_vector_without_gradient():
    sh, cfg = _tiny()
    p = ref.init_params(cfg, 3)
    batch = ref.to_batch(make_batch(sh, seed=5))
    _, _, g, _ = ref.loss_and_grads(p, batch, cfg, 0.0)
    assert float(g["emb_mtx"][0].abs().max()) == 0.0
    p2 = ref.init_params(cfg, 3)
    p2["emb_mtx"][0] = 123.0   # the raw row 0 is masked out of the graph (score.py:45-47)
    y1 = ref.forward(p, batch, cfg)
    y2 = ref.forward(p2, batch, cfg)
    assert torch.equal(y1, y2)


def test_ka3_outputs_invariant_to_ids_beyond_length():
    sh, cfg = _tiny()
    p = ref.init_params(cfg, 3)
    b = make_batch(sh, seed=6, length=3)
    b2 = tuple(x.copy() for x in b)
    rng = np.random.default_rng(0)
    for k in range(4):
        b2[k][:, 3:] = rng.integers(1, sh.feature_size, size=b2[k][:, 3:].shape)
    l1, y1, g1, _ = ref.loss_and_grads(p, ref.to_batch(b), cfg, 1e-4)
    l2, y2, g2, _ = ref.loss_and_grads(p, ref.to_batch(b2), cfg, 1e-4)
    assert torch.equal(y1, y2)
    for n in g1:   # masked slots add exact zeros; only the accumulation order of the dense scatter may differ
        assert torch.allclose(g1[n], g2[n], rtol=1e-5, atol=1e-10), n


def test_ka4_batch_norm_is_a_fixed_affine_map():
    sh, cfg = _tiny()
    p = ref.init_params(cfg, 3)
    p["bn1/gamma"] = torch.rand_like(p["bn1/gamma"]) + 0.5
    p["bn1/beta"] = torch.randn_like(p["bn1/beta"])
    batch = ref.to_batch(make_batch(sh, seed=7))
    y_all = ref.forward(p, batch, cfg)
    one = [x[:1] for x in batch]
    y_one = ref.forward(p, one, cfg)   # no batch statistics: a single sample gives the same prediction
    assert torch.allclose(y_all[:1], y_one, atol=1e-6)
    _, inter = ref.forward(p, batch, cfg, return_intermediates=True)
    x = inter["fc_in"]
    bn = x * p["bn1/gamma"] / math.sqrt(1 + 1e-3) + p["bn1/beta"]
    fc1 = torch.relu(bn @ p["fc1/kernel"] + p["fc1/bias"])
    fc2 = torch.relu(fc1 @ p["fc2/kernel"] + p["fc2/bias"])
    logit = (fc2 @ p["fc3/kernel"] + p["fc3/bias"]).reshape(-1)
    assert torch.allclose(logit, inter["logit"], atol=1e-5)


def test_ka5_ranking_metrics_known_answers():
    n_groups, group = 7, 100
    rng = np.random.default_rng(3)
    preds = rng.permutation(n_groups * group).astype(np.float64).reshape(n_groups, group) / (n_groups * group)
    iids = np.arange(n_groups * group).reshape(n_groups, group) + 1000
    ranks = []
    for g in range(n_groups):
        ranks.append(int((preds[g] > preds[g, 0]).sum()))
    nd5, nd10, hr1, hr5, hr10, mrr = metrics_ref.get_ranking_quality(preds.reshape(-1), iids.reshape(-1))
    assert nd5 == pytest.approx(np.mean([math.log(2) / math.log(p + 2) if p < 5 else 0 for p in ranks]))
    assert nd10 == pytest.approx(np.mean([math.log(2) / math.log(p + 2) if p < 10 else 0 for p in ranks]))
    assert hr1 == pytest.approx(np.mean([p < 1 for p in ranks]))
    assert hr5 == pytest.approx(np.mean([p < 5 for p in ranks]))
    assert hr10 == pytest.approx(np.mean([p < 10 for p in ranks]))
    assert mrr == pytest.approx(np.mean([1.0 / (p + 1) for p in ranks]))


def test_ka6_gru_zero_kernels():
    """zero kernels, gate bias 1, candidate bias 0: u = sigmoid(1), c = 0, h_t = sigmoid(1) * h_{t-1} = 0."""
    B, T, D, H = 2, 5, 6, 4
    x = torch.randn(B, T, D)
    length = torch.tensor([5, 3])
    gk, gb = torch.zeros(D + H, 2 * H), torch.ones(2 * H)
    ck, cb = torch.zeros(D + H, H), torch.full((H,), 0.5)
    out, last = ref.gru_dynamic_rnn(x, length, gk, gb, ck, cb, H)
    u, c = 1 / (1 + math.exp(-1.0)), math.tanh(0.5)
    h = 0.0
    for t in range(T):
        h = u * h + (1 - u) * c
        assert torch.allclose(out[0, t], torch.full((H,), h), atol=1e-6)
        if t < 3:
            assert torch.allclose(out[1, t], torch.full((H,), h), atol=1e-6)
            h3 = h
        else:   # dynamic_rnn: zero output, state copied through
            assert float(out[1, t].abs().max()) == 0.0
    assert torch.allclose(last[1], torch.full((H,), h3), atol=1e-6)


def test_ka7_adam_tf_formulation_and_dense_drift():
    p = {"w": torch.tensor([1.0, -2.0, 0.5])}
    st = ref.AdamState(p)
    g1 = torch.tensor([0.3, -0.02, 1e-3])
    lr = 5e-4
    before = p["w"].clone()
    ref.adam_apply(p, {"w": g1}, st, lr)
    lr_t = lr * math.sqrt(1 - 0.999) / (1 - 0.9)
    assert torch.allclose(st.m["w"], 0.1 * g1, rtol=1e-6)
    assert torch.allclose(st.v["w"], 0.001 * g1 * g1, rtol=1e-5)
    expect = before - lr_t * 0.1 * g1 / (torch.sqrt(0.001 * g1 * g1) + 1e-8)
    assert torch.allclose(p["w"], expect, rtol=1e-6)
    assert torch.allclose(p["w"] - before, -lr * torch.sign(g1), rtol=1e-3)   # ~ -lr*sign(g) when |g| >> 3e-7
    # step 2 with g = 0: slots decay and the variable STILL MOVES (dense-Adam behaviour of score.py:45-47,98)
    mid = p["w"].clone()
    ref.adam_apply(p, {"w": torch.zeros(3)}, st, lr)
    assert torch.allclose(st.m["w"], 0.09 * g1, rtol=1e-5)
    assert torch.allclose(st.v["w"], 0.000999 * g1 * g1, rtol=1e-5)
    assert float((p["w"] - mid).abs().min()) > 1e-5


def test_loss_is_logloss_plus_l2_including_bn():
    sh, cfg = _tiny()
    p = ref.init_params(cfg, 3)
    batch = ref.to_batch(make_batch(sh, seed=8))
    y = ref.forward(p, batch, cfg)
    lam = 0.01
    l2 = sum(float((v.double() ** 2).sum()) / 2 for n, v in p.items()
             if n not in ref.NON_TRAINABLE and "bias" not in n and "emb" not in n)
    assert any(n == "bn1/gamma" for n in p if ref.is_l2_regularised(n))
    yl = batch[6].double()
    ll = float((-yl * torch.log(y.double() + 1e-7) - (1 - yl) * torch.log(1 - y.double() + 1e-7)).mean())
    assert float(ref.total_loss(p, y, batch[6], lam)) == pytest.approx(ll + lam * l2, rel=1e-5)


@pytest.mark.parametrize("model_type", ref.MODEL_TYPES)
def test_fp64_twin_and_finite_differences(model_type):
    sh, cfg = _tiny(model_type)
    p32 = ref.init_params(cfg, 5)
    p64 = ref.init_params(cfg, 5, torch.float64)
    batch = ref.to_batch(make_batch(sh, seed=9, batch=6))
    l32, y32, g32, _ = ref.loss_and_grads(p32, batch, cfg, 1e-3)
    l64, y64, g64, _ = ref.loss_and_grads(p64, batch, cfg, 1e-3)
    assert float(l32) == pytest.approx(float(l64), rel=1e-5)
    assert torch.allclose(y32.double(), y64, atol=1e-5)
    for n in g64:
        scale = float(g64[n].abs().max()) + 1e-12
        assert float((g32[n].double() - g64[n]).abs().max()) <= 2e-4 * scale + 1e-8, n   # dense_5/bias: softmax shift invariance makes its true gradient 0
    # central finite differences in fp64 on a few coordinates of a few variables
    rng = np.random.default_rng(1)
    names = [n for n in g64 if n != "emb_mtx"][:6] + ["fc1/kernel"]
    for n in names:
        flat = p64[n].reshape(-1)
        for _ in range(2):
            i = int(rng.integers(flat.numel()))
            eps = 1e-6
            old = float(flat[i])
            flat[i] = old + eps
            lp = float(ref.total_loss(p64, ref.forward(p64, batch, cfg), batch[6], 1e-3))
            flat[i] = old - eps
            lm = float(ref.total_loss(p64, ref.forward(p64, batch, cfg), batch[6], 1e-3))
            flat[i] = old
            fd = (lp - lm) / (2 * eps)
            assert fd == pytest.approx(float(g64[n].reshape(-1)[i]), rel=1e-4, abs=1e-8), (n, i)


def test_train_returns_pre_update_loss_and_eval_includes_l2():
    sh, _ = _tiny()
    o = ref.ScoreOracle(*sh.ctor_args(), seed=2)
    b = make_batch(sh, seed=10)
    _, _, loss_eval = o.eval(None, b, 1e-3)
    loss_train = o.train(None, b, 5e-4, 1e-3, keep_prob=1.0)
    assert loss_train == pytest.approx(loss_eval, rel=1e-6)
    _, _, loss_after = o.eval(None, b, 1e-3)
    assert loss_after != pytest.approx(loss_eval, rel=1e-9)


def test_loader_dummy_float_rows_are_cast():
    sh, _ = _tiny()
    b = list(make_batch(sh, seed=11, batch=4))
    nested = [x.tolist() for x in b]
    nested[0][0][0] = np.zeros([sh.obj_per_time_slice, sh.item_fnum]).tolist()   # graph_loader.py:90-91
    t = ref.to_batch(nested)
    assert t[0].dtype == torch.int64 and int(t[0][0, 0].abs().max()) == 0
