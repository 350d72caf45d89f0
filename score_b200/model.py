"""Host-side mirror of the reference model interface (code/score/score.py).

``SCORE(feature_size, eb_dim, hidden_size, max_time_len, obj_per_time_slice, user_fnum, item_fnum)``
has the constructor and the ``train / eval / save / restore`` methods of the reference class of the
same name (score.py:188-191, 101-142), so ``train_score.py`` can use it in place of the TensorFlow
graph: ``sess`` is accepted and ignored, ``batch_data`` is the loader's 8-tuple (nested lists, NumPy
arrays, or torch tensors - CUDA tensors are consumed in place).  All arithmetic happens in
libscore_b200.so (hand-written CUDA, sm_100a) through the C ABI of include/score_b200.h.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi


def _load_listfeed():
    """csrc/listfeed.c (CPython extension, host only): built next to the library when missing, like the library itself"""
    try:
        from . import _listfeed
        return _listfeed
    except ImportError:
        pass
    try:
        from . import build as _build
        _build.build_listfeed()
        from . import _listfeed
        return _listfeed
    except Exception:       # no compiler: NumPy's generic list converter does the same job, slower
        return None


_listfeed = _load_listfeed()

MODEL_TYPES = {"SCORE": 0, "RIA": 1, "RCA": 2, "SCORE_USER": 3, "SCORE_ITEM": 4, "RRN": 5}
ADAM_MODES = {"dense": 0, "lazy": 1, "sparse": 2}
TRAIN_KEEP_PROB = 0.8   # score.py:113


class _Batch:
    """Owns the int32 views handed to the C ABI for the duration of one call."""

    def __init__(self, batch_data, cfg):
        if len(batch_data) != 8:
            raise ValueError("batch_data must be the 8-tuple of graph_loader.py:383")
        T, K, fu, fi = cfg["max_time_len"], cfg["obj_per_time_slice"], cfg["user_fnum"], cfg["item_fnum"]
        on_device = all(hasattr(x, "is_cuda") and x.is_cuda for x in batch_data)
        self.keep = []
        ptrs = []
        B = None
        shapes = [(T, K, fi), (T, K, fu), (T, K, fu), (T, K, fi), (fu,), (fi,), (), ()]
        for x, tail in zip(batch_data, shapes):
            if hasattr(x, "data_ptr") and (on_device or not x.is_cuda):
                # torch tensor (CUDA, or CPU incl. pinned): take the pointer, no NumPy round trip
                if str(x.dtype) != "torch.int32" or not x.is_contiguous():
                    import torch
                    x = x.to(torch.int32).contiguous()
                shape = tuple(x.shape)
                ptr = x.data_ptr()
            else:
                if hasattr(x, "is_cuda") and x.is_cuda:      # mixed host / device tuple: stage through the host
                    x = x.cpu().numpy()
                if isinstance(x, (list, tuple)) and _listfeed is not None:
                    # the reference's own feed (graph_loader.py:383): walked with the C API, ~10x NumPy's generic converter
                    a = np.empty((len(x),) + tail, np.int32)
                    try:
                        _listfeed.fill_i32(x, a.shape, a)
                    except (ValueError, TypeError):      # lists of arrays, ragged input: NumPy converts or explains
                        a = np.asarray(x)
                else:
                    a = x if isinstance(x, np.ndarray) else np.asarray(x)
                if a.dtype != np.int32:
                    # nested lists mix ints with the loader's float dummy rows (graph_loader.py:90-91)
                    a = a.astype(np.int32)
                x = a if a.flags.c_contiguous else np.ascontiguousarray(a)
                shape = x.shape
                ptr = x.ctypes.data
            if B is None:
                B = shape[0]
            if tuple(shape) != (B,) + tail:
                raise ValueError("batch tensor has shape %s, expected %s" % (tuple(shape), (B,) + tail))
            self.keep.append(x)
            ptrs.append(ptr)
        self.B = int(B)
        self.struct = _capi.ScoreBatch(*ptrs, self.B, 1 if on_device else 0)


class SCOREBASE(object):
    """One model on one GPU.  Mirrors SCOREBASE (score.py:11-142)."""

    MODEL_TYPE = "SCORE"

    def __init__(self, feature_size, eb_dim, hidden_size, max_time_len, obj_per_time_slice,
                 user_fnum, item_fnum, *, device=0, adam_mode="lazy", seed=1111, init_weights=True,
                 use_graph=True, max_batch=0):
        self._lib = _capi.load()
        self._h = C.c_void_p()
        self.cfg = dict(feature_size=int(feature_size), eb_dim=int(eb_dim), hidden_size=int(hidden_size),
                        max_time_len=int(max_time_len), obj_per_time_slice=int(obj_per_time_slice),
                        user_fnum=int(user_fnum), item_fnum=int(item_fnum))
        self.obj_per_time_slice = obj_per_time_slice   # public attribute of the reference class (score.py:17)
        c = _capi.ScoreConfig(self.cfg["feature_size"], self.cfg["eb_dim"], self.cfg["hidden_size"],
                              self.cfg["max_time_len"], self.cfg["obj_per_time_slice"], self.cfg["user_fnum"],
                              self.cfg["item_fnum"], MODEL_TYPES[self.MODEL_TYPE], ADAM_MODES[adam_mode],
                              int(max_batch), int(seed), 1 if init_weights else 0, 1 if use_graph else 0)
        rc = self._lib.score_create(C.byref(c), int(device), C.byref(self._h))
        if rc != 0:
            msg = self._lib.score_last_error(None)
            msg = msg.decode() if msg else "score_create failed (%d)" % rc
            self._h = C.c_void_p()
            if rc == _capi.ERR_ARG:
                raise ValueError(msg)
            raise RuntimeError(msg)

    # ------------------------------------------------------------------ reference surface
    def train(self, sess, batch_data, lr, reg_lambda, keep_prob=TRAIN_KEEP_PROB):
        """score.py:101-116 -> pre-update loss (float)."""
        b = _Batch(batch_data, self.cfg)
        loss = C.c_float()
        self._check(self._lib.score_train_step(self._h, C.byref(b.struct), lr, reg_lambda, keep_prob, C.byref(loss)))
        return loss.value

    def eval(self, sess, batch_data, reg_lambda):
        """score.py:118-133 -> (preds list, labels list, loss)."""
        b = _Batch(batch_data, self.cfg)
        preds = np.empty(b.B, np.float32)
        loss = C.c_float()
        self._check(self._lib.score_eval(self._h, C.byref(b.struct), reg_lambda, preds.ctypes.data, C.byref(loss)))
        lab = b.keep[6]          # the int32 labels the call consumed (already converted once at the boundary)
        if hasattr(lab, "is_cuda"):
            lab = lab.cpu().numpy()
        return preds.reshape([-1, ]).tolist(), np.asarray(lab).astype(np.int32).reshape([-1, ]).tolist(), loss.value

    def save(self, sess, path):
        """score.py:135-137."""
        self._check(self._lib.score_save(self._h, str(path).encode()))

    def restore(self, sess, path):
        """score.py:139-142."""
        self._check(self._lib.score_restore(self._h, str(path).encode()))
        print('model restored from {}'.format(path))

    # ------------------------------------------------------------------ asynchronous stepping
    def train_async(self, batch_data, lr, reg_lambda, keep_prob=TRAIN_KEEP_PROB):
        b = _Batch(batch_data, self.cfg)
        self._check(self._lib.score_train_step_async(self._h, C.byref(b.struct), lr, reg_lambda, keep_prob))
        self._inflight = b   # keep host arrays alive until wait()

    def wait(self):
        loss = C.c_float()
        self._check(self._lib.score_wait(self._h, C.byref(loss)))
        self._inflight = None
        return loss.value

    # ------------------------------------------------------------------ parity / inspection
    def forward_backward(self, batch_data, reg_lambda, keep_prob=1.0):
        b = _Batch(batch_data, self.cfg)
        loss = C.c_float()
        self._check(self._lib.score_forward_backward(self._h, C.byref(b.struct), reg_lambda, keep_prob, C.byref(loss)))
        return loss.value

    def get_buffer(self, name):
        cnt = C.c_size_t()
        dt = C.c_int()
        self._check(self._lib.score_get_buffer(self._h, name.encode(), None, 0, C.byref(cnt), C.byref(dt)))
        out = np.empty(cnt.value, np.int32 if dt.value == 1 else np.float32)
        self._check(self._lib.score_get_buffer(self._h, name.encode(), out.ctypes.data, out.nbytes, C.byref(cnt), C.byref(dt)))
        return out

    def embedding_row_grads(self):
        """(unique_rows ascending int64, row_grads [U,d]) of the last forward_backward call."""
        heads = self.get_buffer("emb_grad/heads")
        rows = self.get_buffer("emb_grad/seg_rows").reshape(len(heads), self.cfg["eb_dim"])
        sel = heads != 0
        return heads[sel].astype(np.int64), rows[sel]

    def tensor_names(self):
        out = []
        name = C.create_string_buffer(256)
        r, c = C.c_int64(), C.c_int64()
        for i in range(self._lib.score_tensor_count(self._h)):
            self._check(self._lib.score_tensor_info(self._h, i, name, 256, C.byref(r), C.byref(c)))
            out.append((name.value.decode(), (r.value, c.value)))
        return out

    def get_tensor(self, name):
        if name in ("beta1_power", "beta2_power", "step"):
            out = np.empty(1, np.float32)
            self._check(self._lib.score_get_tensor(self._h, name.encode(), out.ctypes.data, 1))
            return out[0]
        base = name
        for suf in ("/Adam_1", "/Adam"):
            if name.endswith(suf):
                base = name[:-len(suf)]
                break
        shape = dict(self.tensor_names())[base]
        out = np.empty(shape[0] * shape[1], np.float32)
        self._check(self._lib.score_get_tensor(self._h, name.encode(), out.ctypes.data, out.size))
        return out.reshape(shape)

    def set_tensor(self, name, value):
        a = np.ascontiguousarray(np.asarray(value, dtype=np.float32).reshape(-1))
        self._check(self._lib.score_set_tensor(self._h, name.encode(), a.ctypes.data, a.size))

    def load_params(self, params):
        """params: mapping TF variable name -> array (e.g. the oracle's init_params)."""
        for name, _ in self.tensor_names():
            v = params[name]
            v = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
            self.set_tensor(name, v)

    def eval_metrics(self, preds, target_iids, labels, group=100):
        """(logloss, auc, ndcg5, ndcg10, hr1, hr5, hr10, mrr): arithmetic of train_score.py:122-163."""
        p = np.ascontiguousarray(np.asarray(preds, np.float32))
        i = np.ascontiguousarray(np.asarray(target_iids, np.int32))
        l = np.ascontiguousarray(np.asarray(labels, np.int32))
        out = np.zeros(9, np.float64)
        self._check(self._lib.score_eval_metrics(self._h, p.ctypes.data, i.ctypes.data, l.ctypes.data, p.size,
                                                 group, out.ctypes.data))
        return tuple(out[:8].tolist())

    # ------------------------------------------------------------------ measurement
    def launch_count(self):
        return int(self._lib.score_launch_count(self._h))

    def enable_probes(self, on=True):
        self._check(self._lib.score_enable_probes(self._h, 1 if on else 0))

    def probe_times(self):
        out = np.zeros(16, np.float64)
        self._lib.score_probe_times(self._h, out.ctypes.data, 16)
        names = ["coatt_fwd", "coatt_bwd", "emb_update", "sort", "step", "fwd_dense", "bwd_dense", "catchup"]
        return {n: (out[2 * i], int(out[2 * i + 1])) for i, n in enumerate(names)}

    def last_step_stats(self):
        out = np.zeros(3, np.int64)
        self._check(self._lib.score_last_step_stats(self._h, out.ctypes.data))
        return dict(positions=int(out[0]), live=int(out[1]), unique_rows=int(out[2]))

    def stream(self):
        s = C.c_void_p()
        self._check(self._lib.score_stream(self._h, C.byref(s)))
        return s.value

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        _capi.check(self._lib, self._h, rc)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.score_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SCORE(SCOREBASE):
    MODEL_TYPE = "SCORE"


class RIA(SCOREBASE):
    MODEL_TYPE = "RIA"


class RCA(SCOREBASE):
    MODEL_TYPE = "RCA"


class SCORE_USER(SCOREBASE):
    MODEL_TYPE = "SCORE_USER"


class SCORE_ITEM(SCOREBASE):
    MODEL_TYPE = "SCORE_ITEM"


class RRN(SCOREBASE):
    """code/slice_models/slice_model.py:155-173 (train_slice.py builds it with the same seven arguments)."""
    MODEL_TYPE = "RRN"
