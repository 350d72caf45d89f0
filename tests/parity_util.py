"""Shared helpers of the GPU parity tests: run the CUDA path and the CPU oracle on identical inputs
and weights and report the error of every intermediate, gradient and updated variable."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import score_ref as ref  # noqa: E402  (tests are allowed to use the oracle as the checker)
from score_b200 import model as sb  # noqa: E402
from score_b200.synth import SHAPES, Shape, make_batch  # noqa: E402

REL_TOL = 1e-5   # BASELINE.json: forward logits and embedding/dense gradients within 1e-5 relative in fp32


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, tiny): relative to the tensor's scale (entries near zero carry the
    absolute rounding error of the largest terms that cancelled)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    scale = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max() / scale)


# Elementwise bar beside the normwise one: |a - b| <= REL_TOL * |b| + ABS_FRAC * max|b| for EVERY entry.  The absolute
# part is the rounding floor of an fp32 sum whose terms are of the tensor's scale (an entry that is the cancelled sum
# of O(100) such terms cannot be reproduced to 1e-5 of ITSELF by any summation order); it is two orders below the
# normwise bar, so entries far below the tensor maximum are checked too.
ABS_FRAC = 1e-6


def grad_abs_frac(n_terms):
    """absolute part of the elementwise bar for a weight gradient that sums `n_terms` rows (B*T of them for the
    per-slice layers): an fp32 sum of n cancelling terms carries ~sqrt(n) * 2^-24 of the TERM scale whatever the order
    (4x head-room: the terms exceed the sum; measured 1.4e-6 on the tiny shape); never looser than the normwise bar"""
    return float(min(REL_TOL, max(ABS_FRAC, 4 * 2.0 ** -24 * np.sqrt(max(n_terms, 1)))))


def elem_excess(a, b, rtol=REL_TOL, abs_frac=ABS_FRAC):
    """max over entries of |a-b| / (rtol*|b| + abs_frac*max|b|): <= 1 means every entry is inside the bar."""
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    if a.size == 0:
        return 0.0
    scale = max(float(np.abs(b).max()), 1e-30)
    return float((np.abs(a - b) / (rtol * np.abs(b) + abs_frac * scale)).max())


class Report(dict):
    """{name: normwise relative error}; add() also records the elementwise excess under 'elem/<name>'."""

    def add(self, name, a, b, scale_floor=None):
        a = np.asarray(a, np.float64)
        b = np.asarray(b, np.float64)
        self[name] = rel_err(a, b)
        self["elem/" + name] = elem_excess(a, b)
        return self[name]


def make_models(shape: Shape, seed=7, adam_mode="dense", model_type="SCORE", use_graph=False, dtype=torch.float32):
    cfg = ref.ScoreConfig(*shape.ctor_args(), model_type=model_type)
    params = ref.init_params(cfg, seed, torch.float32)
    cls = getattr(sb, model_type)
    m = cls(*shape.ctor_args(), adam_mode=adam_mode, init_weights=False, use_graph=use_graph, seed=seed)
    m.load_params(params)
    if dtype != torch.float32:
        params = type(params)((k, v.to(dtype)) for k, v in params.items())
    return cfg, params, m


def forward_backward_report(shape: Shape, batch, seed=7, reg_lambda=1e-4, keep_prob=1.0, model_type="SCORE",
                            dtype=torch.float32, fp64_twin=False):
    """Returns {name: relative error} for every compared tensor plus exact-match flags.
    fp64_twin: also run the oracle's fp64 twin and report, per gradient, 'grad64/<name>' = CUDA vs fp64 and
    'noise/<name>' = the oracle's OWN fp32 result vs fp64 (normwise): how well-conditioned the quantity is in fp32."""
    cfg, params, m = make_models(shape, seed, model_type=model_type, dtype=dtype)
    loss_c = m.forward_backward(batch, reg_lambda, keep_prob)
    B = batch[0].shape[0]
    T, K, H = cfg.max_time_len, cfg.obj_per_time_slice, cfg.hidden_size
    Ds, Dk = cfg.d_side, cfg.d_key
    Dxu, Dxi = cfg.d_side_user, cfg.d_side_item   # GRU input widths per side (= Ds except RRN)
    masks = None
    if keep_prob < 1.0:   # inject the masks the CUDA path drew (a zero output is either dropped or relu-dead)
        g1 = m.get_buffer("fc1").reshape(B, 200)
        g2 = m.get_buffer("fc2").reshape(B, 80)
        masks = (torch.from_numpy((g1 != 0).astype(np.float32)), torch.from_numpy((g2 != 0).astype(np.float32)))
    tb = ref.to_batch(batch)
    loss_o, y_o, g_o, inter = ref.loss_and_grads(params, tb, cfg, reg_lambda, keep_prob, masks)
    rep = Report()
    live = (np.arange(T)[None, :] < np.asarray(batch[7])[:, None])   # [B,T]

    def live_rows(x, width):
        return x.reshape(B, T, width)[live]

    rep["loss"] = rel_err(loss_c, float(loss_o))
    rep.add("y_pred", m.get_buffer("y_pred"), y_o.numpy())
    xu = m.get_buffer("xhg_user").reshape(B * T, Dxu + H)[:, :Dxu]
    xi = m.get_buffer("xhg_item").reshape(B * T, Dxi + H)[:, :Dxi]
    rep.add("user_side", live_rows(xu, Dxu), inter["user_side"].numpy()[live])
    rep.add("item_side", live_rows(xi, Dxi), inter["item_side"].numpy()[live])
    key = m.get_buffer("key").reshape(B * T, Dk)
    rep.add("user_rep_t", key[:, :H].reshape(B, T, H), inter["user_rep_t"].numpy())
    rep.add("item_rep_t", key[:, H:2 * H].reshape(B, T, H), inter["item_rep_t"].numpy())
    if inter.get("atten_info") is not None and model_type != "RIA":
        rep.add("atten_info", live_rows(key[:, 2 * H:], 4 * K), inter["atten_info"].numpy()[live])
    if inter.get("score") is not None:
        rep.add("att_score", m.get_buffer("score").reshape(B, T), inter["score"].numpy().reshape(B, T))
    rep.add("fc_in", m.get_buffer("fc_in").reshape(B, -1), inter["fc_in"].numpy())
    for name, _ in m.tensor_names():
        if name == "emb_mtx" or name in ref.NON_TRAINABLE:
            continue
        a = np.asarray(m.get_buffer("grad/" + name), np.float64)
        b = g_o[name].numpy().reshape(-1).astype(np.float64)
        scale = float(np.abs(b).max())
        if name.endswith("/bias"):
            # a bias gradient is a (often single-element) sum of the same signed terms as its kernel's gradient:
            # its rounding error scales with those terms, not with the cancelled sum -> use the layer's scale
            scale = max(scale, float(g_o[name[:-5] + "/kernel"].abs().max()))
        rep["grad/" + name] = float(np.abs(a - b).max() / max(scale, 1e-30))
        rep["elem/grad/" + name] = float((np.abs(a - b) / (REL_TOL * np.abs(b) + grad_abs_frac(B * T) * max(scale, 1e-30))).max())
    rows_c, vals_c = m.embedding_row_grads()
    rows_o, vals_o = ref.embedding_row_grads(g_o["emb_mtx"])
    if fp64_twin:
        p64 = type(params)((k, v.double()) for k, v in params.items())
        m64 = None if masks is None else tuple(x.double() for x in masks)
        _, _, g_64, _ = ref.loss_and_grads(p64, tb, cfg, reg_lambda, keep_prob, m64)
        for name, _ in m.tensor_names():
            if name == "emb_mtx" or name in ref.NON_TRAINABLE:
                continue
            t = g_64[name].numpy().reshape(-1)
            scale = float(np.abs(t).max())
            if name.endswith("/bias"):
                scale = max(scale, float(g_64[name[:-5] + "/kernel"].abs().max()))
            scale = max(scale, 1e-30)
            rep["grad64/" + name] = float(np.abs(np.asarray(m.get_buffer("grad/" + name), np.float64) - t).max() / scale)
            rep["noise/" + name] = float(np.abs(g_o[name].numpy().reshape(-1).astype(np.float64) - t).max() / scale)
        t = g_64["emb_mtx"].numpy()[rows_o.numpy()]
        scale = max(float(np.abs(t).max()), 1e-30)
        rep["noise/emb_row_grads"] = float(np.abs(vals_o.numpy().astype(np.float64) - t).max() / scale)
        if np.array_equal(rows_c, rows_o.numpy()):
            rep["grad64/emb_row_grads"] = float(np.abs(vals_c.astype(np.float64) - t).max() / scale)
    rep["emb_rows_exact"] = bool(np.array_equal(rows_c, rows_o.numpy()))
    if rep["emb_rows_exact"]:
        rep.add("emb_row_grads", vals_c, vals_o.numpy())
    else:
        common = np.intersect1d(rows_c, rows_o.numpy())
        rep["emb_rows_missing"] = int(len(rows_o) - len(common))
        rep["emb_rows_extra"] = int(len(rows_c) - len(common))
    # gathered ids / neighbor masks: the sanitized key list must equal ids masked by slice liveness
    keys = m.get_buffer("keys")
    rep["keys_exact"] = bool(np.array_equal(keys, expected_keys(batch, cfg)))
    m.close()
    return rep


def expected_keys(batch, cfg):
    """ids in the slice-major position order of embed.cu - per (b,t) slice [user_1hop | item_2hop | user_2hop |
    item_1hop], then target_user, target_item - zeroed where t >= length (masked slices)."""
    T = cfg.max_time_len
    B = np.asarray(batch[0]).shape[0]
    live = (np.arange(T)[None, :] < np.asarray(batch[7])[:, None])
    u1, u2, i1, i2 = (np.asarray(x).astype(np.int32).reshape(B, T, -1) for x in batch[:4])
    if cfg.model_type == "RRN":   # the 2-hop tensors have no consumer (slice_model.py:158-159): their positions carry key 0
        i2, u2 = np.zeros_like(i2), np.zeros_like(u2)
    hist = np.concatenate([u1, i2, u2, i1], axis=2).copy()
    hist[~live] = 0
    return np.concatenate([hist.reshape(-1), np.asarray(batch[4]).astype(np.int32).reshape(-1),
                           np.asarray(batch[5]).astype(np.int32).reshape(-1)])


def train_steps_report(shape: Shape, batches, seed=7, lr=5e-4, reg_lambda=1e-4, adam_mode="dense",
                       use_graph=False, model_type="SCORE"):
    """Run len(batches) training steps (keep_prob=1) on both paths; compare losses and every variable."""
    cfg, params, m = make_models(shape, seed, adam_mode=adam_mode, use_graph=use_graph, model_type=model_type)
    orc = ref.ScoreOracle(*shape.ctor_args(), model_type=model_type, seed=seed)
    rep = {}
    for i, b in enumerate(batches):
        lc = m.train(None, b, lr, reg_lambda, keep_prob=1.0)
        lo = orc.train(None, b, lr, reg_lambda, keep_prob=1.0)
        rep["loss_step%d" % i] = rel_err(lc, lo)
    for name, _ in m.tensor_names():
        rep["var/" + name] = rel_err(m.get_tensor(name), orc.params[name].numpy())
        if name not in ref.NON_TRAINABLE:
            rep["m/" + name] = rel_err(m.get_tensor(name + "/Adam"), orc.opt.m[name].numpy())
            rep["v/" + name] = rel_err(m.get_tensor(name + "/Adam_1"), orc.opt.v[name].numpy())
    m.close()
    return rep


def print_report(title, rep, tol=REL_TOL):
    print("== %s" % title)
    worst = 0.0
    for k, v in rep.items():
        if isinstance(v, bool):
            print("   %-48s %s" % (k, "exact" if v else "MISMATCH"))
        elif isinstance(v, int):
            print("   %-48s %d" % (k, v))
        elif k.startswith("elem/"):
            print("   %-48s %.3f of the elementwise bar%s" % (k, v, "" if v <= 1.0 else "   <-- OUTSIDE"))
        else:
            flag = "" if v <= tol else "   <-- above %.0e" % tol
            worst = max(worst, v)
            print("   %-48s %.3e%s" % (k, v, flag))
    print("   worst relative error: %.3e" % worst)
    return worst
