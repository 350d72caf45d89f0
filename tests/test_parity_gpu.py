"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs and
weights.  Bar (BASELINE.json): gathered ids, neighbor masks and gradient row sets bit-exact; forward
outputs and embedding / dense gradients within 1e-5 relative in fp32."""
import numpy as np
import pytest
import torch

import parity_util as pu
from oracle import score_ref as ref
from score_b200 import model as sb
from score_b200.synth import SHAPES, Shape, make_batch

pytestmark = pytest.mark.gpu

# The attention's last bias feeds a softmax over T: shifting every score leaves the output unchanged,
# so its true gradient is identically 0 and both implementations only produce rounding noise there.
SHIFT_INVARIANT = ("dense_5/bias",)
GRAD_NOISE_ABS = 1e-7


def _check_fb(rep, tol=pu.REL_TOL):
    assert rep["keys_exact"] is True
    assert rep["emb_rows_exact"] is True
    for k, v in rep.items():
        if isinstance(v, bool):
            continue
        assert v <= tol, "%s: relative error %.3e" % (k, v)


@pytest.mark.parametrize("name", ["tiny", "tiny_tb"])
@pytest.mark.parametrize("keep_prob", [1.0, 0.8])
def test_forward_backward_small(name, keep_prob):
    shape = SHAPES[name]
    rep = pu.forward_backward_report(shape, make_batch(shape, seed=11), keep_prob=keep_prob)
    _check_fb(rep)


def test_forward_backward_tmall_shape_batch100():
    """BASELINE.json config 2: Tmall-shape synthetic, K=10, d=16, batch 100 (the reference's own T=11)."""
    shape = SHAPES["tmall"]
    rep = pu.forward_backward_report(shape, make_batch(shape, seed=5))
    _check_fb(rep)


def test_forward_backward_tmall_t10_as_worded():
    shape = SHAPES["tmall_t10"]
    rep = pu.forward_backward_report(shape, make_batch(shape, seed=6, zipf=1.05))
    _check_fb(rep)


def test_forward_backward_k20_long_t_small_table():
    """CCMR-like widths (if=5, uf=1, T=40, K=20) on a small table so the oracle stays fast."""
    shape = Shape("ccmr_small", 60000, 16, 32, 40, 20, 1, 5, 30000, 20000, 8, 38)
    rep = pu.forward_backward_report(shape, make_batch(shape, seed=7))
    _check_fb(rep)


def test_forward_backward_wide_rows_d64_h128():
    """large-vocab widths (d=64, H=128, uf=if=1) on a small table."""
    shape = Shape("lv_small", 50000, 64, 128, 8, 10, 1, 1, 25000, 24000, 16, 6)
    rep = pu.forward_backward_report(shape, make_batch(shape, seed=8))
    _check_fb(rep)


@pytest.mark.parametrize("geom", [(7, 2, 2, 32), (18, 1, 2, 8), (10, 3, 2, 16)])
def test_forward_backward_runtime_geometry_fallback(geom):
    """geometries outside the compiled-in table (embed.cu: SCORE_GEOM_DISPATCH) take the run-time-geometry kernels;
    K > 16 also takes the unpacked softmax path"""
    K, fi, fu, d = geom
    shape = Shape("odd", 30000, d, 32, 5, K, fu, fi, 12000, 15000, 10, 4)
    rep = pu.forward_backward_report(shape, make_batch(shape, seed=21))
    _check_fb(rep)


@pytest.mark.parametrize("model_type", ["RIA", "RCA", "SCORE_USER", "SCORE_ITEM"])
@pytest.mark.parametrize("name", ["tiny", "tiny_tb"])
def test_forward_backward_ablation_classes(model_type, name):
    """score.py:226-369: the four ablation classes run through the same kernels (flags), parity like SCORE."""
    shape = SHAPES[name]
    rep = pu.forward_backward_report(shape, make_batch(shape, seed=31), model_type=model_type)
    _check_fb(rep)


@pytest.mark.parametrize("model_type", ["RIA", "RCA", "SCORE_USER", "SCORE_ITEM"])
def test_train_steps_ablation_classes(model_type):
    shape = SHAPES["tiny"]
    batches = [make_batch(shape, seed=40 + i) for i in range(3)]
    rep = pu.train_steps_report(shape, batches, adam_mode="lazy", model_type=model_type)
    for k, v in rep.items():
        if k.startswith("loss"):
            assert v <= 1e-5, (k, v)


def test_ragged_lengths_dummy_slices_and_single_sample():
    shape = SHAPES["tiny"]
    b = list(make_batch(shape, seed=12, batch=9, dummy_frac=0.4))
    b[7] = np.array([0, 1, 2, 3, 4, 5, 6, 6, 3], np.int32)   # includes length 0 and length == T
    b[0][2] = 0   # a sample whose whole user-1hop history is dummy nodes
    rep = pu.forward_backward_report(shape, tuple(b))
    _check_fb(rep)
    one = tuple(x[:1] for x in make_batch(shape, seed=13, batch=2))
    _check_fb(pu.forward_backward_report(shape, one))


def test_ids_beyond_length_do_not_matter():
    """KA-3: outputs and gradients are invariant to ids at slices t >= length (bit-exact on the GPU path)."""
    shape = SHAPES["tiny"]
    cfg, params, m = pu.make_models(shape)
    b = make_batch(shape, seed=14, length=3)
    b2 = tuple(x.copy() for x in b)
    rng = np.random.default_rng(0)
    for k in range(4):
        b2[k][:, 3:] = rng.integers(1, shape.feature_size, size=b2[k][:, 3:].shape)
    l1 = m.forward_backward(b, 1e-4)
    y1, g1 = m.get_buffer("y_pred"), m.get_buffer("grad/fc1/kernel")
    r1, v1 = m.embedding_row_grads()
    l2 = m.forward_backward(b2, 1e-4)
    assert l1 == l2
    assert np.array_equal(y1, m.get_buffer("y_pred")) and np.array_equal(g1, m.get_buffer("grad/fc1/kernel"))
    r2, v2 = m.embedding_row_grads()
    assert np.array_equal(r1, r2) and np.array_equal(v1, v2)
    m.close()


def test_out_of_range_id_is_an_error():
    shape = SHAPES["tiny"]
    m = sb.SCORE(*shape.ctor_args(), use_graph=False)
    b = list(make_batch(shape, seed=15))
    b[5] = b[5].copy()
    b[5][0, 0] = shape.feature_size   # one past the last row
    with pytest.raises(ValueError, match="outside"):
        m.train(None, tuple(b), 5e-4, 1e-4)
    b[5][0, 0] = -1
    with pytest.raises(ValueError):
        m.eval(None, tuple(b), 1e-4)
    m.close()


def _oracle_adam_on_cuda_grads(shape, batch, lr, reg_lambda, seed=7):
    """Expected post-step state: the oracle's TF-form Adam applied to the gradients the CUDA backward produced."""
    cfg, params, m = pu.make_models(shape, seed)
    m.forward_backward(batch, reg_lambda)
    grads = {}
    for name, shp in m.tensor_names():
        if name in ref.NON_TRAINABLE:
            continue
        if name == "emb_mtx":
            rows, vals = m.embedding_row_grads()
            g = torch.zeros(shp)
            g[torch.from_numpy(rows)] = torch.from_numpy(vals)
        else:
            g = torch.from_numpy(m.get_buffer("grad/" + name).reshape(shp).copy())
        grads[name] = g.reshape(params[name].shape)
    m.close()
    st = ref.AdamState(params)
    ref.adam_apply(params, grads, st, lr)
    return params, st


@pytest.mark.parametrize("mode", ["dense", "lazy"])
def test_optimizer_step_is_bit_exact_given_the_gradients(mode):
    """Sort + segment-reduce + fused row Adam and the dense Adam reproduce TF's ApplyAdam bit for bit,
    including the zero-gradient drift of untouched rows (dense-gradient semantics of score.py:45-47,98)."""
    shape = SHAPES["tiny"]
    lr, lam = 5e-4, 1e-4
    batch = make_batch(shape, seed=21)
    exp_p, exp_st = _oracle_adam_on_cuda_grads(shape, batch, lr, lam)
    cfg, params, m = pu.make_models(shape, adam_mode=mode)
    m.train(None, batch, lr, lam, keep_prob=1.0)
    for name, _ in m.tensor_names():
        assert np.array_equal(m.get_tensor(name), exp_p[name].numpy().reshape(m.get_tensor(name).shape)), name
        if name not in ref.NON_TRAINABLE:
            assert np.array_equal(m.get_tensor(name + "/Adam").reshape(-1), exp_st.m[name].numpy().reshape(-1)), name
            assert np.array_equal(m.get_tensor(name + "/Adam_1").reshape(-1), exp_st.v[name].numpy().reshape(-1)), name
    m.close()


def test_lazy_adam_equals_dense_adam_bitwise_over_many_steps():
    shape = SHAPES["tiny"]
    batches = [make_batch(shape, seed=30 + (i % 4)) for i in range(9)]   # rows are re-touched after gaps
    states = {}
    for mode in ("dense", "lazy"):
        cfg, params, m = pu.make_models(shape, adam_mode=mode)
        losses = [m.train(None, b, 1e-3, 5e-4, keep_prob=1.0) for b in batches]
        _, _, ev = m.eval(None, batches[0], 5e-4)
        states[mode] = (losses, ev, m.get_tensor("emb_mtx"), m.get_tensor("emb_mtx/Adam"), m.get_tensor("emb_mtx/Adam_1"),
                        m.get_tensor("fc1/kernel"))
        m.close()
    a, b = states["dense"], states["lazy"]
    assert a[0] == b[0] and a[1] == b[1]
    for x, y in zip(a[2:], b[2:]):
        assert np.array_equal(x, y)


def test_training_trajectory_tracks_the_oracle():
    shape = SHAPES["tiny_tb"]
    batches = [make_batch(shape, seed=40 + i) for i in range(4)]
    rep = pu.train_steps_report(shape, batches, adam_mode="lazy")
    for k, v in rep.items():
        if k.startswith("loss_step"):
            assert v <= 1e-5, (k, v)
    # Adam divides by sqrt(v): entries whose gradient is rounding noise move by ~lr in a noise-determined
    # direction, so variables are compared at a few-lr absolute scale; slots stay tight.
    assert rep["var/emb_mtx"] <= 1e-5 and rep["m/emb_mtx"] <= 1e-4 and rep["v/emb_mtx"] <= 1e-4
    assert rep["var/fc1/kernel"] <= 1e-4 and rep["var/dense_3/kernel"] <= 1e-3


def test_cuda_graph_replay_equals_direct_launch_bitwise():
    shape = SHAPES["tiny_tb"]
    batches = [make_batch(shape, seed=50 + i) for i in range(5)]
    out = {}
    for graph in (False, True):
        cfg, params, m = pu.make_models(shape, adam_mode="lazy", use_graph=graph)
        losses = [m.train(None, b, 5e-4, 1e-4) for b in batches]   # dropout on: same Philox stream both ways
        out[graph] = (losses, m.get_tensor("emb_mtx"), m.get_tensor("fc2/kernel"))
        m.close()
    assert out[False][0] == out[True][0]
    assert np.array_equal(out[False][1], out[True][1]) and np.array_equal(out[False][2], out[True][2])


def test_eval_matches_oracle_and_reference_return_types():
    shape = SHAPES["tiny"]
    cfg, params, m = pu.make_models(shape)
    b = make_batch(shape, seed=60, neg=5)
    preds, labels, loss = m.eval(None, [x.tolist() for x in b], 1e-3)
    with torch.no_grad():
        y = ref.forward(params, ref.to_batch(b), cfg)
        lo = float(ref.total_loss(params, y, ref.to_batch(b)[6], 1e-3))
    assert isinstance(preds, list) and isinstance(labels, list) and isinstance(loss, float)
    assert labels == b[6].tolist()
    assert pu.rel_err(preds, y.numpy()) <= 1e-5 and abs(loss - lo) <= 1e-5 * abs(lo)
    m.close()


def test_save_restore_roundtrip(tmp_path, capsys):
    shape = SHAPES["tiny"]
    cfg, params, m = pu.make_models(shape, adam_mode="lazy")
    bs = [make_batch(shape, seed=70 + i) for i in range(3)]
    for b in bs[:2]:
        m.train(None, b, 5e-4, 1e-4, keep_prob=1.0)
    path = str(tmp_path / "ckpt")
    m.save(None, path)
    next_loss = m.train(None, bs[2], 5e-4, 1e-4, keep_prob=1.0)
    after = m.get_tensor("emb_mtx")
    m.close()
    m2 = sb.SCORE(*shape.ctor_args(), adam_mode="lazy", init_weights=False, use_graph=False)
    m2.restore(None, path)
    assert "model restored from" in capsys.readouterr().out   # score.py:142
    assert m2.train(None, bs[2], 5e-4, 1e-4, keep_prob=1.0) == next_loss   # Adam slots and step restored too
    assert np.array_equal(m2.get_tensor("emb_mtx"), after)
    with pytest.raises(IOError):
        m2.restore(None, str(tmp_path / "missing"))
    m2.close()


def test_npz_checkpoint_interop_and_log_files(tmp_path):
    """section 8f-3: TF-name-keyed npz checkpoint incl. Adam slots continues training bit-identically; the log /
    result files have train_score.py's names and line formats (train_score.py:86-92, 260-275)"""
    import pickle
    from score_b200 import logs
    shape = SHAPES["tiny"]
    cfg, params, m = pu.make_models(shape, adam_mode="dense")
    bs = [make_batch(shape, seed=90 + i) for i in range(3)]
    for b in bs[:2]:
        m.train(None, b, 5e-4, 1e-4, keep_prob=1.0)
    names = logs.export_npz(m, str(tmp_path / "ckpt.npz"))
    assert "emb_mtx/Adam_1" in names and "beta2_power" in names and "bn1/moving_mean" in names
    assert "bn1/moving_mean/Adam" not in names
    want = m.train(None, bs[2], 5e-4, 1e-4, keep_prob=1.0)
    after = m.get_tensor("fc1/kernel")
    m.close()
    m2 = sb.SCORE(*shape.ctor_args(), adam_mode="dense", init_weights=False, use_graph=False)
    logs.import_npz(m2, str(tmp_path / "ckpt.npz"))
    assert m2.train(None, bs[2], 5e-4, 1e-4, keep_prob=1.0) == want
    assert np.array_equal(m2.get_tensor("fc1/kernel"), after)
    m2.close()
    name = logs.model_name("SCORE", 100, 5e-4, 1e-4)
    assert name == "SCORE_100_0.0005_0.0001"
    assert logs.save_path("tmall", name, root=str(tmp_path)).endswith("save_model_tmall/SCORE_100_0.0005_0.0001/ckpt")
    best = logs.write_train_log("tmall", name, [0.7, 0.6], [0.9, 0.8, 0.7], [0.1, 0.2, 0.15], [0.2, 0.3, 0.25],
                                [0.0, 0.1, 0.0], [0.1, 0.2, 0.1], [0.2, 0.3, 0.2], [0.05, 0.09, 0.07], root=str(tmp_path))
    assert best == 0.09
    with open(tmp_path / "logs_tmall" / (name + ".pkl"), "rb") as f:
        assert len(pickle.load(f)) == 8
    lines = open(tmp_path / "logs_tmall" / (name + ".result")).read().splitlines()
    assert lines[0] == "Result Validation NDCG@5: 0.2" and lines[-1] == "Result Validation MRR: 0.09"
    logs.write_test_result("tmall", name, 10, 0.1, 0.2, 0.0, 0.1, 0.2, 0.05, root=str(tmp_path))
    assert open(tmp_path / "logs_tmall" / (name + "_10.test.result")).read().splitlines()[5] == "Result Test MRR: 0.05"


def test_device_resident_batch_equals_host_batch():
    shape = SHAPES["tiny"]
    cfg, params, m = pu.make_models(shape)
    b = make_batch(shape, seed=80)
    l_host = m.forward_backward(b, 1e-4)
    g_host = m.get_buffer("grad/fc1/kernel")
    dev = tuple(torch.from_numpy(x).cuda() for x in b)
    l_dev = m.forward_backward(dev, 1e-4)
    assert l_host == l_dev and np.array_equal(g_host, m.get_buffer("grad/fc1/kernel"))
    m.close()


def test_full_size_taobao_properties():
    """BASELINE.json config 3 sizes (V = 5 042 754, B = 1024): size-independent properties instead of the oracle."""
    shape = SHAPES["taobao"]
    m = sb.SCORE(*shape.ctor_args(), adam_mode="lazy", use_graph=False, seed=3)
    b = make_batch(shape, seed=90)
    l1 = m.forward_backward(b, 1e-4)
    y1 = m.get_buffer("y_pred")
    rows1, vals1 = m.embedding_row_grads()
    keys = m.get_buffer("keys")
    assert np.array_equal(keys, pu.expected_keys(b, ref.ScoreConfig(*shape.ctor_args())))
    # gradient row set == set of live non-zero ids, ascending, unique  (bit-exact index work)
    assert np.array_equal(rows1, np.unique(keys[keys != 0]))
    spos = m.get_buffer("sorted_pos")
    skeys = m.get_buffer("sorted_keys")
    assert np.all(np.diff(skeys.astype(np.int64)) >= 0)                      # sortedness
    assert np.array_equal(np.sort(spos), np.arange(len(spos)))               # a permutation
    assert np.array_equal(keys[spos], skeys)                                 # payload follows its key
    same = np.diff(skeys.astype(np.int64)) == 0
    assert np.all(np.diff(spos.astype(np.int64))[same] > 0)                  # stable: positions ascend inside a run
    # determinism: a second run is bit-identical (no float atomics anywhere)
    l2 = m.forward_backward(b, 1e-4)
    rows2, vals2 = m.embedding_row_grads()
    assert l1 == l2 and np.array_equal(y1, m.get_buffer("y_pred")) and np.array_equal(vals1, vals2)
    # linearity of the scatter: sum of per-position gradient rows == sum of reduced rows (checksum of checksums)
    gr = m.get_buffer("grad_rows").reshape(-1, shape.eb_dim).astype(np.float64)
    assert np.allclose(gr[keys != 0].sum(0), vals1.astype(np.float64).sum(0), rtol=1e-6, atol=1e-9)
    assert np.isfinite(y1).all() and 0.0 < y1.min() and y1.max() < 1.0
    m.close()


def test_eval_metrics_match_train_score_arithmetic():
    """K-G: logloss / AUC / NDCG / HR / MRR of train_score.py:122-163 on the device."""
    from oracle import metrics_ref
    shape = SHAPES["tiny"]
    m = sb.SCORE(*shape.ctor_args(), use_graph=False)
    rng = np.random.default_rng(5)
    n_groups, group = 44, 100
    preds = rng.random(n_groups * group).astype(np.float32)          # tie-free with overwhelming probability
    assert len(np.unique(preds)) == preds.size
    iids = rng.integers(2000, 4000, size=n_groups * group).astype(np.int32)   # duplicate candidate ids do occur
    labels = np.tile(np.r_[1, np.zeros(group - 1, np.int32)], n_groups).astype(np.int32)
    got = m.eval_metrics(preds, iids, labels, group)
    want = metrics_ref.eval_metrics(preds, labels, iids, group)
    assert got[1] == pytest.approx(want[1], abs=1e-12)              # AUC (bar: 1e-4)
    assert got[0] == pytest.approx(want[0], rel=1e-12)
    for g, w in zip(got[2:], want[2:]):
        assert g == pytest.approx(w, rel=1e-12, abs=1e-15)
    # ties (saturated sigmoids): AUC uses mid-ranks like sklearn; ranks follow the documented stable rule
    preds_t = np.round(preds * 20).astype(np.float32) / 20
    preds_t[::7] = 1.0
    preds_t[3::11] = 0.0
    got = m.eval_metrics(preds_t, iids, labels, group)
    want = metrics_ref.eval_metrics(preds_t, labels, iids, group, stable=True)
    assert got[1] == pytest.approx(want[1], abs=1e-12)
    assert got[0] == pytest.approx(want[0], rel=1e-12)
    for g, w in zip(got[2:], want[2:]):
        assert g == pytest.approx(w, rel=1e-12, abs=1e-15)
    m.close()


# ---------------------------------------------------------------------------------- data-parallel packed exchange
def _dp_layout(n_dense, d, cap):
    dense_off = 128
    keys_off = dense_off + (n_dense + 127) // 128 * 128
    return dense_off, keys_off, keys_off + cap + cap * d


def test_dp_pack_exports_sorted_unique_rows():
    """score_dp_pack: header count, ids ascending, one row per id = the sum of that id's per-position gradient rows."""
    from score_b200 import parallel
    shape = SHAPES["tiny_tb"]
    cfg, params, m = pu.make_models(shape, adam_mode="lazy")
    dp = parallel.DataParallelTrainer(m, 1, 0)
    dp.begin(make_batch(shape, seed=61), 1e-3, 1e-4, keep_prob=1.0)
    cnt = dp.local_count()
    cap = parallel.exchange_capacity([cnt])
    block = dp.pack(cap)
    torch.cuda.synchronize()
    blk = block.cpu().numpy()
    keys = m.get_buffer("keys")
    d = shape.eb_dim
    grad_rows = m.get_buffer("grad_rows").reshape(-1, d)
    uniq = np.unique(keys[keys != 0])
    g_flat = dp._dev("dense_grad", torch.float32).cpu().numpy().copy()   # the flat dense-gradient buffer (aligned offsets)
    n_dense = g_flat.size
    dense_off, keys_off, words = _dp_layout(n_dense, d, cap)
    assert blk.size == words and cnt == len(uniq) and blk[0] == cnt
    got_keys = blk[keys_off:keys_off + cap]
    assert np.array_equal(got_keys[:cnt], uniq.astype(np.int32)) and not got_keys[cnt:].any()
    rows = blk[keys_off + cap:keys_off + cap + cap * d].view(np.float32).reshape(cap, d)[:cnt]
    want = np.zeros((len(uniq), d), np.float64)
    np.add.at(want, np.searchsorted(uniq, keys[keys != 0]), grad_rows[keys != 0].astype(np.float64))
    assert pu.rel_err(rows, want) <= 1e-6
    assert np.array_equal(blk[dense_off:dense_off + n_dense].view(np.float32), g_flat)
    # finishing the step on the single block equals a plain single-GPU step, bit for bit
    loss = dp.finish(block, cap)
    cfg2, params2, m2 = pu.make_models(shape, adam_mode="lazy")
    loss2 = m2.train(None, make_batch(shape, seed=61), 1e-3, 1e-4, keep_prob=1.0)
    assert abs(loss - loss2) <= 1e-6 * abs(loss2)
    for name in ("emb_mtx", "emb_mtx/Adam", "emb_mtx/Adam_1", "fc1/kernel", "dense_3/kernel"):
        assert np.array_equal(m.get_tensor(name), m2.get_tensor(name)), name
    m.close(); m2.close()


@pytest.mark.parametrize("world,mode,graph", [(2, "lazy", True), (3, "dense", False), (8, "lazy", False), (9, "lazy", False)])
def test_dp_packed_exchange_emulated_ranks(world, mode, graph):
    """`world` handles on one GPU driven through the phases of DataParallelTrainer (the all-gather is a torch.cat):
    every replica must end bit-identical, and equal to one model stepping on the concatenated batch."""
    from score_b200 import parallel
    shape = SHAPES["tiny_tb"]
    steps, lr, lam = 4, 1e-3, 1e-4
    per = [[make_batch(shape, seed=700 + 10 * s + r) for r in range(world)] for s in range(steps)]
    glob = [tuple(np.concatenate([b[i] for b in bs], 0) for i in range(8)) for bs in per]
    cfg, params, m1 = pu.make_models(shape, adam_mode=mode)
    l1 = [m1.train(None, g, lr, lam, keep_prob=1.0) for g in glob]
    ms = [pu.make_models(shape, adam_mode=mode, use_graph=graph)[2] for _ in range(world)]
    dps = [parallel.DataParallelTrainer(m, world, r) for r, m in enumerate(ms)]
    l2 = []
    for bs in per:
        for dp, b in zip(dps, bs):
            dp.begin(b, lr, lam, keep_prob=1.0)
        cap = parallel.exchange_capacity([dp.local_count() for dp in dps])
        blocks = [dp.pack(cap) for dp in dps]
        torch.cuda.synchronize()
        gathered = torch.cat(blocks)
        torch.cuda.synchronize()
        losses = [dp.finish(gathered, cap) for dp in dps]
        assert all(x == losses[0] for x in losses)
        l2.append(losses[0])
    assert pu.rel_err(l2, l1) <= 2e-6
    names = ("emb_mtx", "emb_mtx/Adam", "emb_mtx/Adam_1", "fc1/kernel", "dense_3/kernel", "gru_user_side/gru_cell/gates/kernel")
    ref_t = {n: ms[0].get_tensor(n) for n in names}
    for m in ms[1:]:
        for n in names:
            assert np.array_equal(m.get_tensor(n), ref_t[n]), "replicas diverged: " + n
    assert pu.rel_err(ref_t["emb_mtx"], m1.get_tensor("emb_mtx")) <= 2e-4
    assert pu.rel_err(ref_t["emb_mtx/Adam"], m1.get_tensor("emb_mtx/Adam")) <= 2e-4
    assert pu.rel_err(ref_t["emb_mtx/Adam_1"], m1.get_tensor("emb_mtx/Adam_1")) <= 2e-4
    for m in ms + [m1]:
        m.close()
