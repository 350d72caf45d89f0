"""CPU oracle for the SCoRe training / scoring hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *literal* CPU restatement (PyTorch-CPU, fp32 with an fp64 twin via ``dtype``)
of the TensorFlow-1.x graph that ``code/score/score.py`` builds, op for op, including the
things a clean re-implementation would skip: the full-table row-0 mask multiply
(score.py:45-47), the materialised ``[B,T,K,K,3D]`` co-attention concat (score.py:150-156),
the dense ``[V,d]`` embedding gradient and the dense TF-formulation Adam update of every row
every step (score.py:96-99).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module, and only as the checker / the timed CPU baseline.  The product
(``score_b200``) never imports it and has no CPU fallback.

PARITY STATUS.  The reference ships no tests, no golden vectors and no fixtures for this path, and
TensorFlow 1.x cannot be installed in this image (no wheel for Python 3.12, no network), so no TensorFlow-produced vector
of the graph exists.  What pins this restatement instead:
  * the WIRING - which ids are looked up, what is tiled / concatenated / fed to which layer, variable names, shapes and
    creation order, what train() and eval() feed and fetch - against the reference's OWN classes: score.py's SCORE / RIA /
    RCA / SCORE_USER / SCORE_ITEM and slice_model.py's RRN executed unmodified over a stand-in for the TF ops they call
    (tools/tf_shim.py -> tests/golden/refwiring_*.npz -> tests/test_reference_wiring.py: eval, gradients, optimizer steps);
  * the OP SEMANTICS (GRUCell, ApplyAdam, log_loss, batch_normalization, l2_loss) against TensorFlow's published unit-test
    constants (tests/test_tf_known_answers.py);
  * hand-derivable known answers KA-1..KA-7 (tests/test_oracle.py), an fp64 twin and finite-difference gradient checks.
What stays unpinned: TensorFlow's own numerics of the remaining ops (softmax, dense matmul order, sigmoid) at the last bit -
nothing TF-executed can be produced offline.

TF-1.x semantics encoded here (each differs from a PyTorch default):
  * embedding init truncated_normal(0,1) re-drawn beyond 2 sigma (score.py:44);
    dense kernels glorot_uniform, biases zero (tf.layers.dense defaults);
  * GRUCell: [r,u] = sigmoid([x,h] Wg + bg) with bg initialised to 1.0;
    c = tanh([x, r*h] Wc + bc); h' = u*h + (1-u)*c   (score.py:2,205-208);
  * dynamic_rnn(sequence_length): output 0 and state copied through for t >= length;
  * batch_normalization with training=False forever: moving_mean=0, moving_var=1, eps=1e-3
    (score.py:69);
  * log_loss with epsilon 1e-7 inside both logs, mean over the batch (score.py:80);
  * l2_loss = sum(v**2)/2 over trainables whose name has neither 'bias' nor 'emb'
    (score.py:91-94) - this includes bn1/gamma and bn1/beta;
  * Adam, TF kernel form: alpha = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1);
    v += (g*g-v)(1-b2); var -= m*alpha/(sqrt(v)+eps)   (score.py:98);
  * attention padding constant -2**32+1 (score.py:180);
  * sess.run([loss, train_step]) returns the pre-update loss (score.py:102).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np
import torch

MODEL_TYPES = ("SCORE", "RIA", "RCA", "SCORE_USER", "SCORE_ITEM", "RRN")
# RRN is the slice baseline of code/slice_models/slice_model.py:155-173: same constructor,
# placeholders, embedding block (:41-64), build_fc_net (:66-74), build_logloss incl. the L2 term (:76-85), train / eval
# (:111-143) as SCOREBASE; its graph is  user_side = reduce_sum(user_1hop, axis=2), item_side = reduce_sum(item_1hop,
# axis=2) -> two GRUs -> final states -> [user_state || item_state || target_item || target_user] -> fc net.
NO_COATT = ("RCA", "RRN")     # model types without the two co-attention dense layers
NO_ATTENTION = ("RIA", "RRN")  # model types that never call attention()
PAD_VALUE = float(-2 ** 32 + 1)  # score.py:180
BN_EPS = 1e-3                    # tf.layers.batch_normalization default epsilon
LOGLOSS_EPS = 1e-7               # tf.losses.log_loss default epsilon
ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8  # tf.train.AdamOptimizer defaults


@dataclass
class ScoreConfig:
    """Constructor arguments of SCOREBASE (score.py:12-13), same order."""
    feature_size: int
    eb_dim: int
    hidden_size: int
    max_time_len: int
    obj_per_time_slice: int
    user_fnum: int
    item_fnum: int
    model_type: str = "SCORE"

    # derived widths -------------------------------------------------------
    @property
    def d_user(self):  # width of one user node vector
        return self.user_fnum * self.eb_dim

    @property
    def d_item(self):
        return self.item_fnum * self.eb_dim

    @property
    def d_side(self):  # GRU input width: [1hop_seq || 2hop_seq]
        return self.d_user + self.d_item

    @property
    def d_side_user(self):  # GRU input width of the user side (RRN: reduce_sum(user_1hop), slice_model.py:158)
        return self.d_item if self.model_type == "RRN" else self.d_side

    @property
    def d_side_item(self):  # RRN: reduce_sum(item_1hop), slice_model.py:159
        return self.d_user if self.model_type == "RRN" else self.d_side

    @property
    def d_key(self):  # attention key width (score.py:211)
        H, K = self.hidden_size, self.obj_per_time_slice
        if self.model_type in NO_COATT:
            return 2 * H
        return 2 * H + 4 * K

    @property
    def d_fc_in(self):  # width of the prediction-MLP input (score.py:217)
        H = self.hidden_size
        n_state = 1 if self.model_type in ("SCORE_USER", "SCORE_ITEM") else 2
        return n_state * H + self.d_user + self.d_item


def param_specs(cfg: ScoreConfig):
    """TF variable names and shapes in creation order.

    Un-named tf.layers.dense layers are numbered 'dense', 'dense_1', ... in creation order:
    the two co-attention layers first (score.py:196-197), then - after the GRU variables -
    the four layers of attention() (score.py:172-177).  RCA has no co-attention layers, so
    its attention layers start at 'dense'; RIA never calls attention().
    """
    H = cfg.hidden_size
    roles = role_names(cfg)

    def kb(prefix, shape):
        return [(prefix + "/kernel", shape), (prefix + "/bias", (shape[1],))]

    out = [("emb_mtx", (cfg.feature_size, cfg.eb_dim))]
    if cfg.model_type not in NO_COATT:
        out += kb(roles["coatt_item"], (3 * cfg.d_item, 1))
        out += kb(roles["coatt_user"], (3 * cfg.d_user, 1))
    for side, width in (("gru_user_side", cfg.d_side_user), ("gru_item_side", cfg.d_side_item)):
        out += kb(side + "/gru_cell/gates", (width + H, 2 * H))
        out += kb(side + "/gru_cell/candidate", (width + H, H))
    if cfg.model_type not in NO_ATTENTION:
        out += kb(roles["att_q"], (cfg.d_side, cfg.d_key))
        out += kb(roles["att_fc1"], (4 * cfg.d_key, 80))
        out += kb(roles["att_fc2"], (80, 40))
        out += kb(roles["att_fc3"], (40, 1))
    F = cfg.d_fc_in
    out += [("bn1/gamma", (F,)), ("bn1/beta", (F,)),
            ("bn1/moving_mean", (F,)), ("bn1/moving_variance", (F,))]
    out += kb("fc1", (F, 200)) + kb("fc2", (200, 80)) + kb("fc3", (80, 1))
    return out


def role_names(cfg: ScoreConfig):
    """Map logical layer roles to the TF variable prefix for this model type."""
    k = 0
    roles = {}
    if cfg.model_type not in NO_COATT:
        for r in ("coatt_item", "coatt_user"):
            roles[r] = "dense" if k == 0 else "dense_%d" % k
            k += 1
    if cfg.model_type not in NO_ATTENTION:
        for r in ("att_q", "att_fc1", "att_fc2", "att_fc3"):
            roles[r] = "dense" if k == 0 else "dense_%d" % k
            k += 1
    return roles


NON_TRAINABLE = ("bn1/moving_mean", "bn1/moving_variance")


def is_l2_regularised(name: str) -> bool:
    """build_l2norm filter (score.py:92-94)."""
    return name not in NON_TRAINABLE and "bias" not in name and "emb" not in name


def _truncated_normal(shape, gen: np.random.Generator):
    out = gen.standard_normal(shape).astype(np.float32)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = gen.standard_normal(int(bad.sum())).astype(np.float32)
        bad = np.abs(out) > 2.0
    return out


def init_params(cfg: ScoreConfig, seed: int = 1111, dtype=torch.float32):
    """TF-default initialisers (RNG stream is NumPy's, TF's cannot be reproduced)."""
    gen = np.random.default_rng(seed)
    params = OrderedDict()
    for name, shape in param_specs(cfg):
        if name == "emb_mtx":
            val = _truncated_normal(shape, gen)
        elif name.endswith("/kernel"):
            fan_in, fan_out = shape
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            val = gen.uniform(-lim, lim, size=shape).astype(np.float32)
        elif name.endswith("gates/bias"):
            val = np.ones(shape, np.float32)           # GRUCell gate bias init 1.0
        elif name in ("bn1/gamma", "bn1/moving_variance"):
            val = np.ones(shape, np.float32)
        else:
            val = np.zeros(shape, np.float32)
        params[name] = torch.from_numpy(val).to(dtype)
    return params


def to_batch(batch_data):
    """Cast the loader's 8-tuple of nested lists (dummy slices are float zeros,
    graph_loader.py:90-91) to int64 tensors, as feeding int32 placeholders does."""
    out = []
    for x in batch_data:
        a = np.asarray(x)
        out.append(torch.from_numpy(a.astype(np.int64)))
    return out


def _dense(x, kernel, bias, act=None):
    y = torch.matmul(x, kernel) + bias
    if act == "relu":
        y = torch.relu(y)
    return y


def co_attention(seq1, seq2, target_t, kernel, bias):
    """score.py:147-167, literal (both tiles expand on axis 3, so rel[b,t,i,j] depends on i only)."""
    B, T, K, _ = seq1.shape
    target = target_t.unsqueeze(2).unsqueeze(2).expand(B, T, K, K, target_t.shape[-1])
    seq1_tile = seq1.unsqueeze(3).expand(B, T, K, K, seq1.shape[-1])
    seq2_tile = seq2.unsqueeze(3).expand(B, T, K, K, seq2.shape[-1])
    inp = torch.cat([target, seq1_tile, seq2_tile], dim=-1)
    relateness = _dense(inp, kernel, bias, "relu")                       # [B,T,K,K,1]
    atten = torch.softmax(relateness.reshape(B, T, K * K), dim=-1).reshape(B, T, K, K)
    seq1_weights = atten.sum(dim=3).unsqueeze(3)
    seq2_weights = atten.sum(dim=2).unsqueeze(3)
    seq1_result = (seq1 * seq1_weights).sum(dim=2)
    seq2_result = (seq2 * seq2_weights).sum(dim=2)
    relateness = relateness.reshape(B, T, K, K)
    atten_info = torch.cat([relateness.sum(dim=3), relateness.sum(dim=2)], dim=2)
    return seq1_result, seq2_result, atten_info


def gru_dynamic_rnn(x, length, gk, gb, ck, cb, H):
    """tf.nn.dynamic_rnn(GRUCell(H), inputs=x, sequence_length=length) (score.py:205-208)."""
    B, T, _ = x.shape
    h = x.new_zeros(B, H)
    outs = []
    for t in range(T):
        xt = x[:, t]
        value = torch.sigmoid(torch.matmul(torch.cat([xt, h], 1), gk) + gb)
        r, u = value[:, :H], value[:, H:]
        c = torch.tanh(torch.matmul(torch.cat([xt, r * h], 1), ck) + cb)
        new_h = u * h + (1 - u) * c
        alive = (t < length).unsqueeze(1)
        outs.append(torch.where(alive, new_h, torch.zeros_like(new_h)))
        h = torch.where(alive, new_h, h)
    return torch.stack(outs, 1), h


def attention(key, query, mask, p, roles):
    """score.py:169-186: returns softmax weights over T only."""
    B, T, Dk = key.shape
    q = _dense(query, p[roles["att_q"] + "/kernel"], p[roles["att_q"] + "/bias"])
    queries = q.unsqueeze(1).expand(B, T, Dk)
    inp = torch.cat([queries, key, queries - key, queries * key], dim=-1)
    fc1 = _dense(inp, p[roles["att_fc1"] + "/kernel"], p[roles["att_fc1"] + "/bias"], "relu")
    fc2 = _dense(fc1, p[roles["att_fc2"] + "/kernel"], p[roles["att_fc2"] + "/bias"], "relu")
    fc3 = _dense(fc2, p[roles["att_fc3"] + "/kernel"], p[roles["att_fc3"] + "/bias"])
    paddings = torch.ones_like(fc3) * PAD_VALUE
    score = torch.softmax(torch.where(mask == 1, fc3, paddings).reshape(B, T), dim=-1)
    return score.unsqueeze(2)


def forward(params, batch, cfg: ScoreConfig, keep_prob: float = 1.0, dropout_masks=None,
            return_intermediates: bool = False):
    """Forward graph of SCORE / RIA / RCA / SCORE_USER / SCORE_ITEM (score.py:12-369) and RRN (slice_model.py:155-173).

    ``dropout_masks``: optional (mask1 [B,200], mask2 [B,80]) of 0/1 keep flags so a test can
    inject the masks the CUDA path drew; otherwise torch's RNG is used when keep_prob < 1.
    """
    u1_ids, u2_ids, i1_ids, i2_ids, tu_ids, ti_ids, label, length = batch
    p = params
    dtype = p["emb_mtx"].dtype
    roles = role_names(cfg)
    B = u1_ids.shape[0]
    T, K, d, H = cfg.max_time_len, cfg.obj_per_time_slice, cfg.eb_dim, cfg.hidden_size

    # embedding (score.py:43-66): full-table mask multiply, then six lookups
    emb_mask = torch.ones(cfg.feature_size, d, dtype=dtype)
    emb_mask[0] = 0
    emb = p["emb_mtx"] * emb_mask
    user_1hop = emb[u1_ids].reshape(B, T, K, cfg.d_item)
    user_2hop = emb[u2_ids].reshape(B, T, K, cfg.d_user)
    item_1hop = emb[i1_ids].reshape(B, T, K, cfg.d_user)
    item_2hop = emb[i2_ids].reshape(B, T, K, cfg.d_item)
    target_item = emb[ti_ids].reshape(B, cfg.d_item)
    target_user = emb[tu_ids].reshape(B, cfg.d_user)

    mask = (torch.arange(T).unsqueeze(0) < length.unsqueeze(1)).to(dtype).unsqueeze(-1)
    target_user_t = target_user.unsqueeze(1).expand(B, T, cfg.d_user)
    target_item_t = target_item.unsqueeze(1).expand(B, T, cfg.d_item)

    inter = {}
    if cfg.model_type == "RRN":
        # slice_model.py:158-168: the 2-hop lookups exist in the graph (:52-59) but nothing consumes them
        user_side = user_1hop.sum(2)
        item_side = item_1hop.sum(2)
        inter.update(user_side=user_side, item_side=item_side, atten_info=None)
        user_rep_t, user_last = gru_dynamic_rnn(
            user_side, length, p["gru_user_side/gru_cell/gates/kernel"], p["gru_user_side/gru_cell/gates/bias"],
            p["gru_user_side/gru_cell/candidate/kernel"], p["gru_user_side/gru_cell/candidate/bias"], H)
        item_rep_t, item_last = gru_dynamic_rnn(
            item_side, length, p["gru_item_side/gru_cell/gates/kernel"], p["gru_item_side/gru_cell/gates/bias"],
            p["gru_item_side/gru_cell/candidate/kernel"], p["gru_item_side/gru_cell/candidate/bias"], H)
        inter.update(user_rep_t=user_rep_t, item_rep_t=item_rep_t)
        inp = torch.cat([user_last, item_last, target_item, target_user], dim=1)
        return _fc_head(p, inp, inter, keep_prob, dropout_masks, dtype, return_intermediates)
    if cfg.model_type == "RCA":
        user_1hop_seq, user_2hop_seq = user_1hop.sum(2), user_2hop.sum(2)
        item_1hop_seq, item_2hop_seq = item_1hop.sum(2), item_2hop.sum(2)
        atten_info = None
    else:
        user_1hop_seq, item_2hop_seq, info_item = co_attention(
            user_1hop, item_2hop, target_item_t,
            p[roles["coatt_item"] + "/kernel"], p[roles["coatt_item"] + "/bias"])
        user_2hop_seq, item_1hop_seq, info_user = co_attention(
            user_2hop, item_1hop, target_user_t,
            p[roles["coatt_user"] + "/kernel"], p[roles["coatt_user"] + "/bias"])
        if cfg.model_type == "RIA":
            atten_info = info_item + info_user
        else:
            atten_info = torch.cat([info_item, info_user], dim=2)
    user_side = torch.cat([user_1hop_seq, user_2hop_seq], dim=2)
    item_side = torch.cat([item_1hop_seq, item_2hop_seq], dim=2)
    inter.update(user_side=user_side, item_side=item_side, atten_info=atten_info)

    user_rep_t, user_last = gru_dynamic_rnn(
        user_side, length, p["gru_user_side/gru_cell/gates/kernel"], p["gru_user_side/gru_cell/gates/bias"],
        p["gru_user_side/gru_cell/candidate/kernel"], p["gru_user_side/gru_cell/candidate/bias"], H)
    item_rep_t, item_last = gru_dynamic_rnn(
        item_side, length, p["gru_item_side/gru_cell/gates/kernel"], p["gru_item_side/gru_cell/gates/bias"],
        p["gru_item_side/gru_cell/candidate/kernel"], p["gru_item_side/gru_cell/candidate/bias"], H)
    inter.update(user_rep_t=user_rep_t, item_rep_t=item_rep_t)

    if cfg.model_type == "RIA":
        inp = torch.cat([user_last, item_last, target_item, target_user], dim=1)
    else:
        query = torch.cat([target_user, target_item], dim=1)
        if cfg.model_type == "RCA":
            key = torch.cat([user_rep_t, item_rep_t], dim=2)
        else:
            key = torch.cat([user_rep_t, item_rep_t, atten_info], dim=2)
        score = attention(key, query, mask, p, roles)
        inter.update(score=score)
        user_final = (user_rep_t * score).sum(1)
        item_final = (item_rep_t * score).sum(1)
        if cfg.model_type == "SCORE_USER":
            inp = torch.cat([user_final, target_item, target_user], dim=1)
        elif cfg.model_type == "SCORE_ITEM":
            inp = torch.cat([item_final, target_item, target_user], dim=1)
        else:
            inp = torch.cat([user_final, item_final, target_item, target_user], dim=1)

    return _fc_head(p, inp, inter, keep_prob, dropout_masks, dtype, return_intermediates)


def _fc_head(p, inp, inter, keep_prob, dropout_masks, dtype, return_intermediates):
    """build_fc_net (score.py:68-76 = slice_model.py:66-74); BN always in inference mode"""
    inv = p["bn1/gamma"] / torch.sqrt(p["bn1/moving_variance"] + BN_EPS)
    bn1 = inp * inv + (p["bn1/beta"] - p["bn1/moving_mean"] * inv)
    fc1 = _dense(bn1, p["fc1/kernel"], p["fc1/bias"], "relu")
    if keep_prob < 1.0:
        m1 = dropout_masks[0].to(dtype) if dropout_masks is not None else \
            (torch.rand(fc1.shape) < keep_prob).to(dtype)
        fc1 = fc1 * m1 / keep_prob
    fc2 = _dense(fc1, p["fc2/kernel"], p["fc2/bias"], "relu")
    if keep_prob < 1.0:
        m2 = dropout_masks[1].to(dtype) if dropout_masks is not None else \
            (torch.rand(fc2.shape) < keep_prob).to(dtype)
        fc2 = fc2 * m2 / keep_prob
    logit = _dense(fc2, p["fc3/kernel"], p["fc3/bias"]).reshape(-1)
    y_pred = torch.sigmoid(logit)
    inter.update(fc_in=inp, logit=logit)
    if return_intermediates:
        return y_pred, inter
    return y_pred


def total_loss(params, y_pred, label, reg_lambda):
    """build_logloss + build_l2norm (score.py:78-81, 91-94)."""
    y = label.to(y_pred.dtype)
    ll = (-y * torch.log(y_pred + LOGLOSS_EPS) - (1 - y) * torch.log(1 - y_pred + LOGLOSS_EPS)).mean()
    loss = ll
    for name, v in params.items():
        if is_l2_regularised(name):
            loss = loss + reg_lambda * (v * v).sum() / 2
    return loss


def loss_and_grads(params, batch, cfg, reg_lambda, keep_prob=1.0, dropout_masks=None):
    """Pre-update loss, predictions and the dense gradient of every trainable variable."""
    leaves = OrderedDict()
    for n, v in params.items():
        leaves[n] = v.detach().clone().requires_grad_(n not in NON_TRAINABLE)
    y_pred, inter = forward(leaves, batch, cfg, keep_prob, dropout_masks, return_intermediates=True)
    loss = total_loss(leaves, y_pred, batch[6], reg_lambda)
    names = [n for n in leaves if n not in NON_TRAINABLE]
    grads = torch.autograd.grad(loss, [leaves[n] for n in names], allow_unused=True)
    gdict = OrderedDict()
    for n, g in zip(names, grads):
        gdict[n] = torch.zeros_like(leaves[n]) if g is None else g
    return loss.detach(), y_pred.detach(), gdict, {k: (v.detach() if v is not None else None)
                                                  for k, v in inter.items()}


def embedding_row_grads(emb_grad: torch.Tensor):
    """(unique_rows int64 ascending, row_grads [U,d]) of a dense embedding gradient: the rows a
    sparse scatter must touch.  Row 0 never appears (its gradient is masked to zero)."""
    touched = (emb_grad != 0).any(dim=1)
    rows = torch.nonzero(touched).reshape(-1)
    return rows, emb_grad[rows]


class AdamState:
    """tf.train.AdamOptimizer slots: <var>/Adam (m), <var>/Adam_1 (v), beta1_power, beta2_power."""

    def __init__(self, params):
        self.m = OrderedDict((n, torch.zeros_like(v)) for n, v in params.items() if n not in NON_TRAINABLE)
        self.v = OrderedDict((n, torch.zeros_like(v)) for n, v in params.items() if n not in NON_TRAINABLE)
        dt = next(iter(params.values())).dtype
        self.beta1_power = torch.tensor(ADAM_B1, dtype=dt)
        self.beta2_power = torch.tensor(ADAM_B2, dtype=dt)


def adam_apply(params, grads, st: AdamState, lr: float):
    """Dense ApplyAdam on every trainable variable, in place, in the op order of TF's kernel:
        alpha = lr * sqrt(1 - beta2_power) / (1 - beta1_power)
        m += (g - m) * (1 - beta1);  v += (g*g - v) * (1 - beta2);  var -= (m * alpha) / (sqrt(v) + eps)
    Evaluated with NumPy element-wise ops in the variables' dtype: every step is one correctly rounded
    IEEE operation (Eigen's SSE/AVX sqrt and div are too), which torch's CPU sqrt (a vendor vector-math
    routine) does not guarantee to the last bit."""
    np_dt = np.float32 if st.beta1_power.dtype == torch.float32 else np.float64
    f = np_dt
    b1p, b2p = f(st.beta1_power.item()), f(st.beta2_power.item())
    alpha = f(f(f(lr) * np.sqrt(f(1) - b2p, dtype=np_dt)) / f(f(1) - b1p))
    omb1, omb2, eps = f(f(1) - f(ADAM_B1)), f(f(1) - f(ADAM_B2)), f(ADAM_EPS)
    for n, g in grads.items():
        gn = g.detach().numpy().astype(np_dt, copy=False)
        m, v, var = st.m[n].numpy(), st.v[n].numpy(), params[n].numpy()   # views: updated in place
        m += (gn - m) * omb1
        v += (gn * gn - v) * omb2
        var -= (m * alpha) / (np.sqrt(v) + eps)
    st.beta1_power = torch.tensor(f(b1p * f(ADAM_B1)), dtype=st.beta1_power.dtype)
    st.beta2_power = torch.tensor(f(b2p * f(ADAM_B2)), dtype=st.beta2_power.dtype)


class ScoreOracle:
    """Same surface as the reference model classes (score.py:101-142); ``sess`` is ignored."""

    def __init__(self, feature_size, eb_dim, hidden_size, max_time_len, obj_per_time_slice,
                 user_fnum, item_fnum, model_type="SCORE", seed=1111, dtype=torch.float32):
        self.cfg = ScoreConfig(feature_size, eb_dim, hidden_size, max_time_len,
                               obj_per_time_slice, user_fnum, item_fnum, model_type)
        self.params = init_params(self.cfg, seed, dtype)
        self.opt = AdamState(self.params)

    def train(self, sess, batch_data, lr, reg_lambda, keep_prob=0.8, dropout_masks=None):
        batch = to_batch(batch_data)
        loss, _, grads, _ = loss_and_grads(self.params, batch, self.cfg, reg_lambda, keep_prob, dropout_masks)
        adam_apply(self.params, grads, self.opt, lr)
        return float(loss)

    def eval(self, sess, batch_data, reg_lambda):
        batch = to_batch(batch_data)
        with torch.no_grad():
            y_pred = forward(self.params, batch, self.cfg, 1.0)
            loss = total_loss(self.params, y_pred, batch[6], reg_lambda)
        return y_pred.reshape(-1).tolist(), batch[6].reshape(-1).tolist(), float(loss)
