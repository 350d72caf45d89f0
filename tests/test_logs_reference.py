"""SURVEY.md section 8 f-3, pinned to the reference: train_score.py's OWN train() / restore() / eval() functions (lifted
out of the file with ast, executed unmodified) drive a stand-in model through the interface the drop-in keeps -
SCORE(...), .train, .eval, .save, .restore - and write their checkpoint path, logs_<ds>/<name>.pkl, <name>.result and
<name>_<K>.test.result; score_b200.logs must produce the same names and the same bytes from the same numbers.

Build container only (reads /root/reference); nothing here needs a GPU - the model is a deterministic stub, what is
under test is the file / naming contract around it."""
import ast
import os
import pickle

import numpy as np
import pytest

from score_b200 import logs

REF = "/root/reference/code/score/train_score.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="needs the reference checkout (build container only)")


class StubModel(object):
    """the interface train_score.py uses (score.py:101-142); numbers are deterministic functions of the call count"""
    saved, restored, created = [], [], []

    def __init__(self, *args):
        StubModel.created.append(args)
        self.n_train = 0
        self.n_eval = 0

    def train(self, sess, batch_data, lr, reg_lambda):
        self.n_train += 1
        return 0.7 / (1.0 + 0.1 * self.n_train)

    def eval(self, sess, batch_data, reg_lambda):
        self.n_eval += 1
        rng = np.random.default_rng(1000 + self.n_eval + 17 * self.n_train)
        n = len(batch_data[6])
        preds = rng.random(n) * 0.5
        preds[::100] += 0.3 * min(1.0, self.n_train / 6.0)       # the positives climb as training proceeds
        return preds.tolist(), [int(x) for x in np.asarray(batch_data[6]).reshape(-1)], 0.69 - 0.01 * self.n_train

    def save(self, sess, path):
        StubModel.saved.append(path)

    def restore(self, sess, path):
        StubModel.restored.append(path)


class StubLoader(object):
    """GraphLoader(graph_handler_params, batch_size, target_file, start_time, pred_time, worker_n, neg_sample_num)"""

    def __init__(self, params, batch_size, target_file, start_time, pred_time, worker_n, neg):
        self.batch_size, self.neg, self.n = batch_size, neg, 4 if neg == 1 else 3

    def __iter__(self):
        for b in range(self.n):
            B = self.batch_size
            items = [[1000 + (b * B + i), 7] for i in range(B)]
            labels = [1 if i % (1 + self.neg) == 0 else 0 for i in range(B)]
            yield [None, None, None, None, [[1]] * B, items, labels, [3] * B]


class _Ctx(object):
    def __init__(self, **kw):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def run(self, *a, **k):
        return None


class StubTF(object):
    GPUOptions = ConfigProto = Session = _Ctx
    global_variables_initializer = local_variables_initializer = staticmethod(lambda: None)


def reference_functions():
    """every function of train_score.py + its upper-case module constants, compiled from the reference's own source"""
    tree = ast.parse(open(REF).read())
    keep = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            keep.append(node)
        elif isinstance(node, ast.Assign) and all(isinstance(t, ast.Name) and t.id.isupper() for t in node.targets):
            keep.append(node)
    import math
    import time
    from sklearn.metrics import log_loss, roc_auc_score
    ns = {"os": os, "np": np, "pkl": pickle, "math": math, "time": time, "log_loss": log_loss, "roc_auc_score": roc_auc_score,
          "tf": StubTF, "GraphLoader": StubLoader}
    for cls in ("SCORE", "RIA", "RCA", "SCORE_USER", "SCORE_ITEM"):
        ns[cls] = StubModel
    exec(compile(ast.Module(body=keep, type_ignores=[]), REF, "exec"), ns)
    return ns


def test_train_and_restore_files_match_the_references_own_functions(tmp_path, monkeypatch, capsys):
    ns = reference_functions()
    ref_dir, our_dir = tmp_path / "ref", tmp_path / "ours"
    ref_dir.mkdir(); our_dir.mkdir()
    monkeypatch.chdir(ref_dir)
    StubModel.saved, StubModel.restored, StubModel.created = [], [], []
    model_type, bs, lr, reg, K = "SCORE", 4, 5e-4, 1e-4, 10
    ctor = (5000, 16, 32, 11, K, 3, 4)
    best = ns["train"]("tmall", "train.txt", "vali.txt", None, 0, 9, 10, model_type, bs, ctor[0], ctor[1], ctor[2], ctor[3], K, lr,
                       reg, 12, None, None, 3, 4)
    ns["restore"]("tmall", "test.txt", None, 0, 11, model_type, bs, ctor[0], ctor[1], ctor[2], ctor[3], K, lr, reg, None, None, 3, 4)
    out = capsys.readouterr().out
    assert "STEP 0  LOSS TRAIN: NULL" in out and "RESTORE, LOSS TEST:" in out
    # the constructor call the drop-in must accept (score.py:188-191, positional)
    assert StubModel.created and all(c == ctor for c in StubModel.created)

    name = logs.model_name(model_type, bs, lr, reg)
    ref_logs = ref_dir / "logs_tmall"
    assert sorted(os.listdir(ref_logs)) == sorted([name + ".pkl", name + ".result", "%s_%d.test.result" % (name, K)])
    # checkpoint path: what train() passed to model.save and restore() to model.restore
    assert StubModel.saved, "the stub's validation MRR never improved: no checkpoint was written"
    want_ckpt = os.path.normpath(os.path.relpath(logs.save_path("tmall", name, root=str(our_dir)), str(our_dir)))
    assert {os.path.normpath(p) for p in StubModel.saved} == {want_ckpt}
    assert [os.path.normpath(p) for p in StubModel.restored] == [want_ckpt]
    assert os.path.isdir(ref_dir / "save_model_tmall" / name)

    # the training log: same 8-tuple -> same bytes, same .result text, same return value
    with open(ref_logs / (name + ".pkl"), "rb") as f:
        curves = pickle.load(f)
    assert len(curves) == 8 and len(curves[1]) == len(curves[7]) >= 2 and len(curves[0]) == len(curves[1]) - 1
    ours_best = logs.write_train_log("tmall", name, *curves, root=str(our_dir))
    assert ours_best == best
    our_logs = our_dir / "logs_tmall"
    for fn in (name + ".pkl", name + ".result"):
        assert open(our_logs / fn, "rb").read() == open(ref_logs / fn, "rb").read(), fn

    # the test result file: the numbers restore() printed are the ones the reference's eval() computed from the stub
    lines = open(ref_logs / ("%s_%d.test.result" % (name, K))).read().splitlines()
    vals = [float(l.split(": ")[1]) for l in lines]
    logs.write_test_result("tmall", name, K, *[np.float64(v) for v in vals], root=str(our_dir))
    assert open(our_logs / ("%s_%d.test.result" % (name, K))).read() == open(ref_logs / ("%s_%d.test.result" % (name, K))).read()


def test_references_eval_equals_the_metric_oracle_on_the_same_predictions(monkeypatch, tmp_path):
    """eval() of train_score.py:144-163 run unmodified over the stub loader / model vs oracle/metrics_ref.py on the same
    predictions (the CUDA metric kernel is held to that oracle and to the reference-produced fixture in tests/test_golden.py)"""
    from oracle import metrics_ref
    ns = reference_functions()
    m = StubModel()
    got = ns["eval"](m, None, None, "vali.txt", 0, 10, 1e-4)
    m2 = StubModel()
    preds, labels, iids, losses = [], [], [], []
    for batch in StubLoader(None, 100, "vali.txt", 0, 10, 8, 99):
        p, l, loss = m2.eval(None, batch, 1e-4)
        preds += p; labels += l; losses.append(loss)
        iids += [r[0] for r in batch[5]]
    want = metrics_ref.eval_metrics(np.asarray(preds), np.asarray(labels), np.asarray(iids), group=100)
    np.testing.assert_allclose(got[:8], want[:8], rtol=1e-12, atol=1e-15)
    assert got[8] == pytest.approx(sum(losses) / len(losses), rel=1e-15)


class FakeStore(object):
    """stands in for GraphStore.sample on a CPU box: same signature, host tensors, ids derived from the arguments"""
    user_fnum, item_fnum = 1, 2
    calls = []

    def sample(self, uids, iids, group, start_time, pred_time, max_time_len, obj_per_time_slice, mode="rs", seed=0,
               draw_id=0, stream=None, as_numpy=False):
        import torch
        B, T, K = len(iids), max_time_len, obj_per_time_slice
        FakeStore.calls.append((B, group, start_time, pred_time, T, K, mode, seed, draw_id))
        z = lambda f: torch.zeros(B, T, K, f, dtype=torch.int32)
        tu = torch.as_tensor(np.repeat(np.asarray(uids, np.int32), group)[:B].reshape(B, 1))
        ti = torch.as_tensor(np.stack([np.asarray(iids, np.int32), np.full(B, 7, np.int32)], 1))
        label = torch.as_tensor((np.arange(B) % group == 0).astype(np.int32))
        return (z(2), z(1), z(1), z(2), tu, ti, label, torch.full((B,), pred_time - start_time, dtype=torch.int32))


def test_reference_signature_graph_loader_serves_the_references_own_eval_and_train(tmp_path, monkeypatch, capsys):
    """score_b200.graph.GraphLoader takes the reference's constructor arguments (graph_loader.py:279-281, the
    graph_handler_params list of train_score.py:292-297) and feeds train_score.py's OWN eval() / train() unchanged:
    batch count incl. the short last batch, one uid per 1 + neg items, labels, lengths, and np.array(batch_data[5])[:, 0]
    (train_score.py:157) on a tensor that otherwise lives in device memory.  The sampling itself is the CUDA sampler's
    (GPU tests); a host stand-in for GraphStore.sample lets this run without a device."""
    from score_b200 import graph
    ns = reference_functions()
    ns["GraphLoader"] = graph.GraphLoader
    graph.register_graph("tmall_2hop", FakeStore())
    params = [12, "tmall_2hop", 10, 50, 80, 0, 200, 500, "rs", None, None, 1, 2]
    target = tmp_path / "target_10_sample.txt"
    rng = np.random.default_rng(3)
    n_users = 5                                     # 5 lines x 100 items: EVAL_BATCH_SIZE 100 -> 5 batches of one line
    target.write_text("".join("%d,%s\n" % (u + 1, ",".join(str(51 + int(x)) for x in rng.integers(0, 80, 100))) for u in range(n_users)))
    FakeStore.calls = []
    m = StubModel()
    monkeypatch.chdir(tmp_path)
    out = ns["eval"](m, None, params, str(target), 0, 10, 1e-4)
    assert len(out) == 9 and m.n_eval == n_users
    assert all(c[:7] == (100, 100, 0, 10, 11, 10, "rs") for c in FakeStore.calls)          # T = 12 - 0 - 1, group 1 + 99
    # training: batch 4 with 1 negative -> 2 lines per batch, 5 lines -> 3 batches per epoch, the last one short
    train_file = tmp_path / "target_9.txt"
    train_file.write_text(target.read_text())
    FakeStore.calls = []
    StubModel.saved, StubModel.restored, StubModel.created = [], [], []
    ns["train"]("tmall", str(train_file), str(target), params, 0, 9, 10, "SCORE", 4, 5000, 16, 32, 11, 10, 5e-4, 1e-4, 12,
                None, None, 1, 2)
    train_calls = [c for c in FakeStore.calls if c[1] == 2]
    assert train_calls and {c[0] for c in train_calls} == {4, 2} and all(c[3] == 9 and c[4] == 11 for c in train_calls)
    assert len({c[7] for c in train_calls}) > 1                                             # a new loader (epoch) draws anew
    # the reference's own error behaviour for a batch size that is not a multiple of 1 + neg (graph_loader.py:289-291)
    with pytest.raises(SystemExit):
        graph.GraphLoader(params, 5, str(train_file), 0, 9, 8, 1)
    assert "batch size should be time of 2" in capsys.readouterr().out
    with pytest.raises(KeyError):
        graph.GraphLoader([12, "unknown_db"] + params[2:], 4, str(train_file), 0, 9, 8, 1)
    # the target-item tensor reads like the reference's nested list AND like a device tensor
    b = next(iter(graph.GraphLoader(params, 4, str(train_file), 0, 9, 8, 1)))
    assert np.array(b[5])[:, 0].tolist() == [int(x) for x in train_file.read_text().splitlines()[0].split(",")[1:3]] + \
        [int(x) for x in train_file.read_text().splitlines()[1].split(",")[1:3]]
    assert hasattr(b[5], "data_ptr") and b[5].is_cuda is False and tuple(b[5].shape) == (4, 2)
    # and the model boundary takes the batch as it is (pointer path of _Batch: no NumPy round trip)
    from score_b200 import model as sb
    bb = sb._Batch(b, dict(max_time_len=11, obj_per_time_slice=10, user_fnum=1, item_fnum=2))
    assert bb.B == 4 and bb.struct.target_item == b[5].data_ptr()
