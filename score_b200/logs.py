"""Checkpoint interop and the log / result files of train_score.py.

* ``export_npz / import_npz``: every global variable of the model keyed by its TensorFlow name - ``emb_mtx``,
  ``dense/kernel`` ... ``fc3/bias``, the Adam slots ``<var>/Adam`` / ``<var>/Adam_1`` and ``beta1_power`` /
  ``beta2_power`` - i.e. what ``tf.train.Saver`` stores for this graph (score.py:135-137).  A TF checkpoint converts
  with ``np.savez(path, **{n: reader.get_tensor(n) for n in reader.get_variable_to_shape_map()})``.
* ``model_name / save_path / write_train_log / write_test_result``: the directory layout and file contents of
  train_score.py:86-92, 248-275, so the downstream scripts that read ``logs_<ds>/*.pkl`` / ``*.result`` keep working.
"""
from __future__ import annotations

import os
import pickle as pkl

import numpy as np

NON_TRAINABLE = ("bn1/moving_mean", "bn1/moving_variance")


def export_npz(model, path):
    out = {}
    for name, _ in model.tensor_names():
        out[name] = model.get_tensor(name)
        if name not in NON_TRAINABLE:
            out[name + "/Adam"] = model.get_tensor(name + "/Adam")
            out[name + "/Adam_1"] = model.get_tensor(name + "/Adam_1")
    out["beta1_power"] = np.float32(model.get_tensor("beta1_power"))
    out["beta2_power"] = np.float32(model.get_tensor("beta2_power"))
    out["step"] = np.int64(model.get_tensor("step"))   # not a TF variable: the exact optimizer step (see import_npz)
    np.savez(path, **out)
    return sorted(out)


def import_npz(model, path, strict=True):
    data = np.load(path)
    names = [n for n, _ in model.tensor_names()]
    missing = [n for n in names if n not in data.files]
    if missing and strict:
        raise KeyError("checkpoint lacks variables: %s" % ", ".join(missing))
    for name in names:
        if name not in data.files:
            continue
        model.set_tensor(name, data[name])
        for suf in ("/Adam", "/Adam_1"):
            if name not in NON_TRAINABLE and name + suf in data.files:
                model.set_tensor(name + suf, data[name + suf])
    # The optimizer step is recovered from the slot variables (beta1_power = 0.9^(step+1); beta2_power once that has gone
    # denormal) unless the file carries it explicitly; either way every imported row counts as current at that step
    # (lazy Adam: nothing is replayed against the imported state).
    for p in ("beta1_power", "beta2_power"):
        if p in data.files:
            model.set_tensor(p, np.asarray([data[p]], np.float32))
    if "step" in data.files:
        model.set_tensor("step", np.asarray([data["step"]], np.float32))


def model_name(model_type, train_batch_size, lr, reg_lambda):
    return '{}_{}_{}_{}'.format(model_type, train_batch_size, lr, reg_lambda)          # train_score.py:248


def save_path(data_set, name, root="."):
    d = os.path.join(root, 'save_model_{}/{}/'.format(data_set, name))                  # train_score.py:249-251
    if not os.path.exists(d):
        os.makedirs(d)
    return os.path.join(d, 'ckpt')


def write_train_log(data_set, name, train_losses, vali_losses, vali_ndcgs_5, vali_ndcgs_10, vali_hrs_1, vali_hrs_5,
                    vali_hrs_10, vali_mrrs, root="."):
    """train_score.py:260-275 -> best validation MRR (the value train() returns)"""
    d = os.path.join(root, 'logs_{}/'.format(data_set))
    if not os.path.exists(d):
        os.makedirs(d)
    with open(os.path.join(d, '{}.pkl'.format(name)), 'wb') as f:
        pkl.dump((train_losses, vali_losses, vali_ndcgs_5, vali_ndcgs_10, vali_hrs_1, vali_hrs_5, vali_hrs_10, vali_mrrs), f)
    index = int(np.argmax(vali_mrrs))
    with open(os.path.join(d, '{}.result'.format(name)), 'w') as f:
        f.write('Result Validation NDCG@5: {}\n'.format(vali_ndcgs_5[index]))
        f.write('Result Validation NDCG@10: {}\n'.format(vali_ndcgs_10[index]))
        f.write('Result Validation HR@1: {}\n'.format(vali_hrs_1[index]))
        f.write('Result Validation HR@5: {}\n'.format(vali_hrs_5[index]))
        f.write('Result Validation HR@10: {}\n'.format(vali_hrs_10[index]))
        f.write('Result Validation MRR: {}\n'.format(vali_mrrs[index]))
    return vali_mrrs[index]


def write_test_result(data_set, name, obj_per_time_slice, ndcg_5, ndcg_10, hr_1, hr_5, hr_10, mrr, root="."):
    """train_score.py:86-92"""
    d = os.path.join(root, 'logs_{}/'.format(data_set))
    if not os.path.exists(d):
        os.makedirs(d)
    with open(os.path.join(d, '{}_{}.test.result'.format(name, obj_per_time_slice)), 'w') as f:
        f.write('Result Test NDCG@5: {}\n'.format(ndcg_5))
        f.write('Result Test NDCG@10: {}\n'.format(ndcg_10))
        f.write('Result Test HR@1: {}\n'.format(hr_1))
        f.write('Result Test HR@5: {}\n'.format(hr_5))
        f.write('Result Test HR@10: {}\n'.format(hr_10))
        f.write('Result Test MRR: {}\n'.format(mrr))
