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
        if k.startswith("elem/"):
            # every entry: |a-b| <= 1e-5 |b| + 1e-6 max|b| (parity_util.elem_excess); the shift-invariant bias holds noise only
            assert v <= 1.0 or k.endswith(SHIFT_INVARIANT), "%s: %.3f x the elementwise bar" % (k, v)
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


@pytest.mark.parametrize("model_type", ["RIA", "RCA", "SCORE_USER", "SCORE_ITEM", "RRN"])
@pytest.mark.parametrize("name", ["tiny", "tiny_tb"])
def test_forward_backward_ablation_classes(model_type, name):
    """score.py:226-369: the four ablation classes, and RRN (slice_models/slice_model.py:155-173: 1-hop sum pooling,
    per-side GRU widths, final states), run through the same kernels (flags), parity like SCORE."""
    shape = SHAPES[name]
    rep = pu.forward_backward_report(shape, make_batch(shape, seed=31), model_type=model_type)
    _check_fb(rep)


@pytest.mark.parametrize("model_type", ["RIA", "RCA", "SCORE_USER", "SCORE_ITEM", "RRN"])
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


# ---------------------------------------------------------------------------------- TF-published known answers on the CUDA path
def _ka_model():
    """d=8, H=32, T=2, K=2, uf=if=1 on a 16-row table, every variable zero except what a test sets."""
    V, d, H, T, K = 16, 8, 32, 2, 2
    m = sb.SCORE(V, d, H, T, K, 1, 1, init_weights=False, use_graph=False, adam_mode="dense")
    shapes = dict(m.tensor_names())
    z = {n: np.zeros(s, np.float32) for n, s in shapes.items()}
    z["bn1/gamma"][:] = 1.0
    z["bn1/moving_variance"][:] = 1.0
    for side in ("gru_user_side", "gru_item_side"):
        z[side + "/gru_cell/gates/bias"][:] = 1.0     # GRUCell's own initial value
    return m, z, (V, d, H, T, K)


def _ka_batch(B, T, K, u1_ids):
    """user_1hop[b,t,:,0] = u1_ids[b][t] for both neighbors; every other id is the dummy 0"""
    u1 = np.zeros((B, T, K, 1), np.int32)
    for b in range(B):
        for t in range(T):
            u1[b, t, :, 0] = u1_ids[b][t]
    zi = np.zeros((B, T, K, 1), np.int32)
    return [u1, zi.copy(), zi.copy(), zi.copy(), np.zeros((B, 1), np.int32), np.zeros((B, 1), np.int32),
            np.zeros(B, np.int32), np.full(B, T, np.int32)]


def test_tf_known_answer_gru_cell_on_the_cuda_path():
    """rnn_cell_test.py::testGRUCell (kernels 0.5, gate bias 1, candidate bias 0): x=[1,1], h=[0.1,0.1] -> 0.175991 and
    x=[1,1,1], h=[0.1,0.1] -> 0.156736, reproduced by gru_fwd_kernel through the C ABI.  dynamic_rnn starts at h=0, so
    step 1 is crafted to leave exactly 0.1 in the state (update gate sigmoid(0), candidate tanh(atanh 0.2)) through an
    input column that is 0 in step 2; state rows 2.. carry weight 0, so units 0,1 see the published 2-unit cell
    (tests/test_tf_known_answers.py runs the same construction through the oracle)."""
    import math
    m, z, (V, d, H, T, K) = _ka_model()
    Ds = 2 * d
    emb = np.zeros((V, d), np.float32)
    emb[1, 3] = 1.0                  # step-1 trigger row
    emb[2, :2] = 1.0                 # x = [1, 1]
    emb[3, :3] = 1.0                 # x = [1, 1, 1]
    gk = z["gru_user_side/gru_cell/gates/kernel"]        # [Ds + H, 2H]
    ck = z["gru_user_side/gru_cell/candidate/kernel"]    # [Ds + H, H]
    gk[:3] = 0.5; gk[Ds:Ds + 2] = 0.5; ck[:3] = 0.5; ck[Ds:Ds + 2] = 0.5
    gk[3, H:] = -1.0
    ck[3, :] = math.atanh(0.2)
    z["emb_mtx"] = emb
    m.load_params(z)
    batch = _ka_batch(2, T, K, [[1, 2], [1, 3]])
    m.eval(None, batch, 0.0)
    key = m.get_buffer("key").reshape(2 * T, -1)[:, :H]   # user_rep_t
    assert np.allclose(key[0], 0.1, atol=1e-7) and np.allclose(key[2], 0.1, atol=1e-7)
    assert np.allclose(key[1], 0.175991, rtol=1e-6, atol=1e-6)      # assertAllClose defaults of the TF test
    assert np.allclose(key[3], 0.156736, rtol=1e-6, atol=1e-6)
    m.close()


def test_tf_known_answer_log_loss_on_the_cuda_path():
    """losses_test.py::LogLossTest: predictions [.9 .2 .2 .8 .4 .6], labels [1 0 1 1 0 0], epsilon 1e-7 ->
    -sum(l log(p+eps) + (1-l) log(1-p+eps)) / 6.  The prediction head is wired as the identity on target_item[0]
    (relu split / re-join, BN scale exactly 1), which holds logit(p)."""
    m, z, (V, d, H, T, K) = _ka_model()
    p = np.asarray([.9, .2, .2, .8, .4, .6])
    lab = np.asarray([1, 0, 1, 1, 0, 0], np.int32)
    emb = np.zeros((V, d), np.float32)
    emb[1:7, 0] = np.log(p / (1 - p))
    z["emb_mtx"] = emb
    z["bn1/moving_variance"][:] = np.float32(1.0) - np.float32(1e-3)   # var + eps == 1.0f exactly
    col = 2 * H                                                        # fc_in = [user_final | item_final | target_item | target_user]
    z["fc1/kernel"][col, 0] = 1.0; z["fc1/kernel"][col, 1] = -1.0
    z["fc2/kernel"][0, 0] = 1.0; z["fc2/kernel"][1, 1] = 1.0
    z["fc3/kernel"][0, 0] = 1.0; z["fc3/kernel"][1, 0] = -1.0
    m.load_params(z)
    batch = _ka_batch(6, T, K, [[0, 0]] * 6)
    batch[5] = np.arange(1, 7, dtype=np.int32).reshape(6, 1)           # target_item
    batch[6] = lab
    preds, labels, loss = m.eval(None, batch, 0.0)
    eps = 1e-7
    want = -np.sum(lab * np.log(p + eps) + (1 - lab) * np.log(1 - p + eps)) / 6.0
    assert np.allclose(preds, p, rtol=1e-6, atol=1e-7)
    assert loss == pytest.approx(want, rel=2e-6)
    # testAllCorrectNoLossWeight: saturated correct predictions give ~0 (the epsilon keeps the logs finite)
    emb[1:7, 0] = np.where(lab == 1, 40.0, -40.0)
    m.set_tensor("emb_mtx", emb)
    _, _, loss0 = m.eval(None, batch, 0.0)
    assert loss0 == pytest.approx(0.0, abs=1e-3)
    m.close()


def test_tf_known_answer_adam_update_numpy_on_the_cuda_path():
    """adam_test.py::adam_update_numpy (TF's own NumPy reference of ApplyAdam), applied in fp64 to the gradients the
    CUDA backward reports, against the variables / slots the CUDA optimizer leaves - dense variables and embedding
    rows, three steps (as testBasic), learning rate 0.001."""
    from test_tf_known_answers import adam_update_numpy
    shape = SHAPES["tiny"]
    cfg, params, m = pu.make_models(shape, adam_mode="dense")
    names = ["fc1/kernel", "gru_item_side/gru_cell/gates/bias", "dense/kernel", "emb_mtx"]
    var = {n: m.get_tensor(n).astype(np.float64) for n in names}
    mm = {n: np.zeros_like(var[n]) for n in names}
    vv = {n: np.zeros_like(var[n]) for n in names}
    lam = 1e-4
    for t in range(1, 4):
        b = make_batch(shape, seed=300 + t)
        m.forward_backward(b, lam)                 # gradients of the full loss (L2 term included), no update
        g = {}
        for n in names[:-1]:
            g[n] = m.get_buffer("grad/" + n).reshape(var[n].shape).astype(np.float64)
        rows, vals = m.embedding_row_grads()
        ge = np.zeros_like(var["emb_mtx"])
        ge[rows] = vals
        g["emb_mtx"] = ge
        m.train(None, b, 0.001, lam, keep_prob=1.0)
        for n in names:
            var[n], mm[n], vv[n] = adam_update_numpy(var[n], g[n], t, mm[n], vv[n])
            assert np.allclose(m.get_tensor(n), var[n], rtol=2e-6, atol=2e-7), (n, t)
            # m mixes gradients of both signs: an entry that cancels carries the fp32 rounding of its terms
            assert np.allclose(m.get_tensor(n + "/Adam"), mm[n], rtol=1e-5, atol=1e-6 * np.abs(mm[n]).max()), (n, t)
            # ApplyAdam computes (1 - beta2) in the variable's type: 1 - 0.999f = 0.00099998713, 1.29e-5 below the fp64
            # 0.001 of the NumPy reference (TF's own adam_test compares the variables only) -> 3e-5 on the v slot
            assert np.allclose(m.get_tensor(n + "/Adam_1"), vv[n], rtol=3e-5, atol=1e-15), (n, t)
    assert float(m.get_tensor("beta1_power")) == pytest.approx(0.9 ** 4, rel=1e-6)
    assert float(m.get_tensor("beta2_power")) == pytest.approx(0.999 ** 4, rel=1e-6)
    m.close()


def test_full_size_taobao_values_match_the_oracle():
    """BASELINE.json config 3 at its real size (V = 5 042 754, B = 1024): loss, predictions, every intermediate, dense
    gradients and the embedding row gradients against the oracle (one oracle step is ~1 s on the box's host cores)."""
    shape = SHAPES["taobao"]
    rep = pu.forward_backward_report(shape, make_batch(shape, seed=91), fp64_twin=True)
    pu.print_report("taobao full size", rep)
    # With N(0,1) rows, 1024 samples and the attention over large atten_info values, some gradients are ill-conditioned
    # in fp32: the oracle's OWN fp32 result is 1e-4 .. 7e-3 away from its fp64 twin ('noise/...', measured: co-attention
    # kernels 1.6e-3, attention MLP 3e-3 .. 7e-3, GRU 5e-5 .. 1e-4, embedding rows 2.4e-3), and two fp32 evaluations with
    # different (valid) summation orders differ by as much.  The bar for a gradient is therefore: 1e-5 against the fp32
    # oracle, OR as close to the fp64 result as the reference's own fp32 arithmetic gets (factor 1.5).
    noise = {k[len("noise/"):]: v for k, v in rep.items() if k.startswith("noise/")}
    g64 = {k[len("grad64/"):]: v for k, v in rep.items() if k.startswith("grad64/")}
    strict = {}
    for k, v in rep.items():
        if k.startswith(("noise/", "grad64/")):
            continue
        base = k[len("elem/"):] if k.startswith("elem/") else k
        name = base[len("grad/"):] if base.startswith("grad/") else base
        if name in g64 and not isinstance(v, bool):
            ok32 = v <= (1.0 if k.startswith("elem/") else pu.REL_TOL)
            ok64 = g64[name] <= max(pu.REL_TOL, 1.5 * noise[name])
            assert ok32 or ok64, "%s: %.3e vs fp32 oracle, %.3e vs fp64 (oracle fp32 noise %.3e)" % (k, v, g64[name], noise[name])
            continue
        strict[k] = v
    _check_fb(strict)


# ---------------------------------------------------------------------------------- ADVICE.md round 1
def test_lazy_adam_window_rolls_over(monkeypatch):
    """a short alpha window (SCORE_ALPHA_CAP) forces several re-bases inside 14 steps: LAZY stays bit-identical to DENSE
    (tf.train.AdamOptimizer has no step limit)"""
    shape = SHAPES["tiny"]
    batches = [make_batch(shape, seed=30 + (i % 4)) for i in range(14)]
    states = {}
    for mode in ("dense", "lazy"):
        if mode == "lazy":
            monkeypatch.setenv("SCORE_ALPHA_CAP", "6")
        cfg, params, m = pu.make_models(shape, adam_mode=mode, use_graph=(mode == "lazy"))
        losses = [m.train(None, b, 1e-3, 5e-4, keep_prob=1.0) for b in batches]
        states[mode] = (losses, m.get_tensor("emb_mtx"), m.get_tensor("emb_mtx/Adam"), m.get_tensor("emb_mtx/Adam_1"))
        m.close()
    assert states["dense"][0] == states["lazy"][0]
    for x, y in zip(states["dense"][1:], states["lazy"][1:]):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("with_step", [True, False])
def test_npz_import_in_lazy_mode_keeps_the_optimizer_state(tmp_path, with_step):
    """import_npz into a LAZY model: the imported rows are current at the imported step (nothing is replayed against
    them), with the explicit step or with the step recovered from beta1_power"""
    from score_b200 import logs
    shape = SHAPES["tiny"]
    cfg, params, m = pu.make_models(shape, adam_mode="lazy")
    bs = [make_batch(shape, seed=90 + i) for i in range(5)]
    for b in bs[:3]:
        m.train(None, b, 5e-4, 1e-4, keep_prob=1.0)
    path = str(tmp_path / "ckpt.npz")
    logs.export_npz(m, path)
    if not with_step:
        d = dict(np.load(path)); d.pop("step"); np.savez(path, **d)
    want = [m.train(None, b, 5e-4, 1e-4, keep_prob=1.0) for b in bs[3:]]
    after = (m.get_tensor("emb_mtx"), m.get_tensor("emb_mtx/Adam"), m.get_tensor("emb_mtx/Adam_1"))
    m.close()
    m2 = sb.SCORE(*shape.ctor_args(), adam_mode="lazy", init_weights=False, use_graph=False)
    logs.import_npz(m2, path)
    assert int(m2.get_tensor("step")) == 3
    assert [m2.train(None, b, 5e-4, 1e-4, keep_prob=1.0) for b in bs[3:]] == want
    for x, y in zip(after, (m2.get_tensor("emb_mtx"), m2.get_tensor("emb_mtx/Adam"), m2.get_tensor("emb_mtx/Adam_1"))):
        assert np.array_equal(x, y)
    # a slot value TF produces late in training: beta1_power underflows to 0 near step 985 - the step then comes from beta2_power
    m2.set_tensor("beta2_power", np.asarray([0.999 ** 2001], np.float32))
    m2.set_tensor("beta1_power", np.asarray([0.0], np.float32))
    assert abs(int(m2.get_tensor("step")) - 2000) <= 1
    m2.close()


def test_dp_exchange_buffers_regrow_between_steps():
    """the packed block and the sampled lists are re-allocated when a later step needs more room (a
    stale pointer was freed twice): step 1 small batch, step 2 a 12x larger one, replicas stay identical to one model
    stepping on the concatenated batches"""
    from score_b200 import parallel
    shape = SHAPES["tiny_tb"]
    world, lr, lam = 2, 1e-3, 1e-4
    sizes = [8, 96, 16, 200]
    per = [[make_batch(shape, seed=800 + 10 * s + r, batch=n) for r in range(world)] for s, n in enumerate(sizes)]
    glob = [tuple(np.concatenate([b[i] for b in bs], 0) for i in range(8)) for bs in per]
    cfg, params, m1 = pu.make_models(shape, adam_mode="lazy")
    l1 = [m1.train(None, g, lr, lam, keep_prob=1.0) for g in glob]
    ms = [pu.make_models(shape, adam_mode="lazy")[2] for _ in range(world)]
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
    for n in ("emb_mtx", "emb_mtx/Adam_1", "fc1/kernel"):
        assert np.array_equal(ms[0].get_tensor(n), ms[1].get_tensor(n)), n
    assert pu.rel_err(ms[0].get_tensor("emb_mtx"), m1.get_tensor("emb_mtx")) <= 2e-4
    for m in ms + [m1]:
        m.close()


@pytest.mark.parametrize("world", [1, 3, 8])
def test_shard_plan_kernel_groups_positions_by_owner(world):
    """score_shard_plan (CUDA) against the bucketing rule of parallel.ExchangePlan stated in NumPy: owner = id % world,
    position order inside a group (stable), owner-local row = id // world + 1, dummy positions last and keyless."""
    import ctypes as C
    from score_b200 import _capi
    shape = SHAPES["tiny_tb"]
    m = sb.SCORE(*shape.ctor_args(), init_weights=False, use_graph=False)
    batch = make_batch(shape, seed=77, dummy_frac=0.3)
    b = sb._Batch(batch, m.cfg)
    m._check(m._lib.score_prepare_batch(m._h, C.byref(b.struct)))
    keys = m.get_buffer("keys").astype(np.int64)
    plan = _capi.ScoreShardPlan()
    m._check(m._lib.score_shard_plan(m._h, world, C.byref(plan)))
    n = int(plan.n_positions)
    assert n == keys.size
    torch.cuda.synchronize()   # the plan kernels run on the handle's stream

    def dev_i32(ptr, cnt):
        out = np.empty(cnt, np.int32)
        m._check(m._lib.score_copy_to_host(out.ctypes.data, ptr, out.nbytes))
        return out

    counts = dev_i32(plan.counts, world + 1)
    owner = np.where(keys != 0, keys % world, world)
    assert np.array_equal(counts, np.bincount(owner, minlength=world + 1))
    order = np.argsort(owner, kind="stable")
    n_valid = n - counts[world]
    assert np.array_equal(dev_i32(plan.send_rows, n_valid), keys[order[:n_valid]] // world + 1)
    mini = dev_i32(plan.mini_keys, n)
    want_mini = np.zeros(n, np.int64)
    want_mini[order[:n_valid]] = np.arange(1, n_valid + 1)
    assert np.array_equal(mini, want_mini)
    m.close()


def test_lazy_never_touched_mark_equals_dense():
    """device-initialised tables keep last_step = -1 for rows no optimizer step has touched (nothing to replay: their
    slots are zero); LAZY with that mark must stay bit-identical to DENSE, rows re-touched after gaps included, and a
    row written from outside (set_rows) loses the mark."""
    shape = SHAPES["tiny_tb"]
    batches = [make_batch(shape, seed=60 + (i % 3)) for i in range(7)]
    out = {}
    for mode in ("dense", "lazy"):
        m = sb.SCORE(*shape.ctor_args(), adam_mode=mode, init_weights=True, seed=77, use_graph=(mode == "lazy"))
        losses = [m.train(None, b, 1e-3, 1e-4, keep_prob=1.0) for b in batches[:4]]
        # rewrite the slots of a few rows from outside in the middle of the run (same values in both modes)
        rows = np.full((5, shape.eb_dim), 0.01, np.float32)
        m._check(m._lib.score_set_rows(m._h, b"emb_mtx/Adam", 10, 5, rows.ctypes.data))
        losses += [m.train(None, b, 1e-3, 1e-4, keep_prob=1.0) for b in batches[4:]]
        out[mode] = (losses, m.get_tensor("emb_mtx"), m.get_tensor("emb_mtx/Adam"), m.get_tensor("emb_mtx/Adam_1"))
        m.close()
    assert out["dense"][0] == out["lazy"][0]
    for a, b in zip(out["dense"][1:], out["lazy"][1:]):
        assert np.array_equal(a, b)


def test_hot_row_longer_than_a_chunk():
    """one side-feature id on every item (a run of ~10 K positions in the sorted list, 5 chunks of RUN_CHUNK = 2048,
    scatter.cu): the chunked long-run reduce of the scatter kernel against the oracle, gradient rows and three optimizer
    steps, LAZY == DENSE bitwise"""
    shape = SHAPES["tiny_tb"]
    hot = shape.feature_size - 1

    def batch(seed):
        b = [x.copy() for x in make_batch(shape, seed=seed)]
        for k in (0, 3, 5):                      # user_1hop, item_2hop, target_item carry item nodes: field 1 = side feature
            live = b[k][..., 0] != 0
            b[k][..., 1] = np.where(live, hot, 0)
        return tuple(b)

    rep = pu.forward_backward_report(shape, batch(5))
    # the hot row is a cancelling sum of ~10 K gradient rows: its entries carry ~sqrt(10^4) * 2^-24 of the TERM scale in any
    # summation order, above the 1e-6 absolute part of the elementwise bar (the normwise 1e-5 bar stays)
    assert rep.pop("elem/emb_row_grads") <= 4.0
    _check_fb(rep)
    keys = np.concatenate([np.asarray(x).reshape(-1) for x in batch(5)[:6]])
    assert (keys == hot).sum() > 3 * 2048
    out = {}
    for mode in ("dense", "lazy"):
        cfg, params, m = pu.make_models(shape, adam_mode=mode)
        losses = [m.train(None, batch(50 + i), 1e-3, 1e-4, keep_prob=1.0) for i in range(3)]
        out[mode] = (losses, m.get_tensor("emb_mtx"), m.get_tensor("emb_mtx/Adam_1"))
        m.close()
    assert out["dense"][0] == out["lazy"][0]
    assert np.array_equal(out["dense"][1], out["lazy"][1]) and np.array_equal(out["dense"][2], out["lazy"][2])
    orc = ref.ScoreOracle(*shape.ctor_args(), seed=7)
    lo = [orc.train(None, batch(50 + i), 1e-3, 1e-4, keep_prob=1.0) for i in range(3)]
    assert pu.rel_err(out["lazy"][0], lo) <= 1e-5
    assert pu.rel_err(out["lazy"][1][hot], orc.params["emb_mtx"].numpy()[hot]) <= 1e-5


@pytest.mark.parametrize("world,mode", [(3, "lazy"), (8, "lazy"), (2, "dense")])
def test_sharded_peer_push_emulated_ranks(world, mode):
    """`world` handles on ONE GPU driven through the phases of the row-sharded step with the PEER-MEMORY kernels
    (score_shard_serve_push / score_shard_grad_push: every "peer" buffer is a tensor of this process, the barrier is a
    device synchronize, the id exchange a host-side regrouping): losses and owned table rows must equal one model
    stepping on the concatenated batch - this pins the offset arithmetic of shard.cu for world sizes a small box cannot run."""
    import ctypes as C
    from score_b200 import _capi, parallel
    shape = SHAPES["tiny_tb"]
    d = shape.eb_dim
    steps, lr, lam = 3, 1e-3, 1e-4
    per = [[make_batch(shape, seed=800 + 10 * s + r, batch=16) for r in range(world)] for s in range(steps)]
    glob = [tuple(np.concatenate([b[i] for b in bs], 0) for i in range(8)) for bs in per]
    cfg, params, m1 = pu.make_models(shape, adam_mode=mode)
    l1 = [m1.train(None, g, lr, lam, keep_prob=1.0) for g in glob]
    ms = []
    for r in range(world):
        a = list(shape.ctor_args())
        a[0] = parallel.shard_rows(shape.feature_size, world)
        m = sb.SCORE(*a, adam_mode=mode, init_weights=False, use_graph=False)
        for name, _ in m.tensor_names():
            v = params[name]
            if name == "emb_mtx":
                v = parallel.global_to_local_table(v, world, r)
            m.set_tensor(name, v.numpy())
        ms.append(m)
    lib = ms[0]._lib
    dev = torch.device("cuda", 0)
    W = world

    def view(ptr, n, dt):
        return torch.as_tensor(parallel._DevView(ptr, (int(n),), "<f4" if dt == torch.float32 else "<i4"), device=dev)

    l2 = []
    for bs in per:
        gb = sum(b[0].shape[0] for b in bs)
        plans, keep = [], []
        for r, (m, b) in enumerate(zip(ms, bs)):
            bb = sb._Batch(b, m.cfg); keep.append(bb)
            m._check(lib.score_set_sample_offset(m._h, sum(x[0].shape[0] for x in bs[:r])))
            m._check(lib.score_prepare_batch(m._h, C.byref(bb.struct)))
            p = _capi.ScoreShardPlan()
            m._check(lib.score_shard_plan(m._h, W, C.byref(p)))
            plans.append(p)
        torch.cuda.synchronize()
        cm = torch.stack([view(p.counts, W + 1, torch.int32).clone() for p in plans]).contiguous()      # [W, W+1] on the device
        cmh = cm.cpu().numpy()
        n_valid = [int(cmh[r, :W].sum()) for r in range(W)]
        n_recv = [int(cmh[:, o].sum()) for o in range(W)]
        sends = [view(plans[r].send_rows, max(n_valid[r], 1), torch.int32)[:n_valid[r]].cpu().numpy() for r in range(W)]
        so = np.concatenate([np.zeros((W, 1), np.int64), np.cumsum(cmh[:, :W], axis=1)], axis=1)        # send_off(r, o)
        wants = [torch.from_numpy(np.concatenate([sends[r][so[r, o]:so[r, o + 1]] for r in range(W)]).astype(np.int32)).to(dev)
                 for o in range(W)]                                                                     # the id all-to-all
        n_pos = int(plans[0].n_positions)
        staged = [torch.zeros((n_pos + 1) * d, dtype=torch.float32, device=dev) for _ in range(W)]
        owned = [torch.zeros(max(n_recv[o], 1) * d, dtype=torch.float32, device=dev) for o in range(W)]
        peers_staged = (C.c_uint64 * W)(*[t.data_ptr() for t in staged])
        peers_owned = (C.c_uint64 * W)(*[t.data_ptr() for t in owned])
        for o, m in enumerate(ms):
            m._check(lib.score_shard_serve_push(m._h, wants[o].data_ptr(), n_recv[o], cm.data_ptr(), W, o, peers_staged))
        torch.cuda.synchronize()
        for r, m in enumerate(ms):
            m._check(lib.score_step_begin(m._h, None, lr, lam, 1.0, gb, 1, staged[r].data_ptr(), plans[r].mini_keys))
        torch.cuda.synchronize()
        g = [view(*_devbuf(m, "dense_grad"), torch.float32) for m in ms]
        tot = torch.stack(g).sum(0)
        for x in g:
            x.copy_(tot)
        for r, m in enumerate(ms):
            m._check(lib.score_shard_grad_push(m._h, cm.data_ptr(), W, r, peers_owned))
        torch.cuda.synchronize()
        data = 0.0
        for o, m in enumerate(ms):
            loss2 = (C.c_float * 2)()
            m._check(lib.score_step_finish(m._h, wants[o].data_ptr(), owned[o].data_ptr(), n_recv[o], loss2))
            data += loss2[0] - loss2[1]
            l2_term = loss2[1]
        l2.append(data + l2_term)
    assert pu.rel_err(l2, l1) <= 2e-6
    emb1 = m1.get_tensor("emb_mtx")
    for r, m in enumerate(ms):
        loc = parallel.global_to_local_table(torch.from_numpy(emb1), world, r).numpy()
        assert pu.rel_err(m.get_tensor("emb_mtx")[1:], loc[1:]) <= 2e-6, r
        assert pu.rel_err(m.get_tensor("fc1/kernel"), m1.get_tensor("fc1/kernel")) <= 2e-6
    for m in ms + [m1]:
        m.close()


def _devbuf(m, name):
    import ctypes as C
    ptr, cnt = C.c_void_p(), C.c_size_t()
    m._check(m._lib.score_device_buffer(m._h, name.encode(), C.byref(ptr), C.byref(cnt)))
    return ptr.value, cnt.value
