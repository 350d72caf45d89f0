"""Host-side mirror of the reference loader on top of the on-GPU graph store (csrc/sampler.cu).

``GraphStore`` holds what ``GraphHandler`` reads from MongoDB (graph_loader.py:52-92: per-node documents with
'1hop' / '2hop' / 'degrees' lists per time slice, plus the side-feature dicts) as CSR arrays in HBM;
``DeviceGraphLoader`` iterates a target file like ``GraphLoader`` (graph_loader.py:279-403) and yields 8-tuples whose
id tensors already live in device memory, so ``model.train(sess, batch, lr, reg)`` consumes them without a copy.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi

MODES = {"rs": 0, "is": 1}


def docs_to_csr(user_docs, item_docs, n_user, n_item, n_slices):
    """user_docs / item_docs: mappings id -> {'1hop': [S lists], '2hop': [S lists], 'degrees': [S lists]}
    (the reference's Mongo documents, graph_storage.py:186-195).  Missing nodes get empty lists."""
    n_nodes = n_user + n_item + 1
    off1 = np.zeros(n_nodes * n_slices + 1, np.int64)
    off2 = np.zeros(n_nodes * n_slices + 1, np.int64)
    ids1, ids2, deg2 = [], [], []
    for node in range(n_nodes):
        doc = user_docs.get(node) if node <= n_user else item_docs.get(node)
        for s in range(n_slices):
            h1 = doc["1hop"][s] if doc is not None else []
            h2 = doc["2hop"][s] if doc is not None else []
            dg = doc["degrees"][s] if doc is not None and "degrees" in doc else [2] * len(h2)
            ids1.extend(h1); ids2.extend(h2); deg2.extend(dg)
            off1[node * n_slices + s + 1] = len(ids1)
            off2[node * n_slices + s + 1] = len(ids2)
    return (off1, np.asarray(ids1, np.int32), off2, np.asarray(ids2, np.int32), np.asarray(deg2, np.int32))


def feat_table(feat_dict, lo, n, width):
    """feat_dict[str(id)] -> list of `width` ints (pickled dicts of feateng_*.py) as a dense [(n+1), width] table
    with row id - lo + 1 ... i.e. row r holds node lo - 1 + r; row 0 is the dummy."""
    out = np.zeros((n + 1, width), np.int32)
    if feat_dict is None or width == 0:
        return out
    for k, v in feat_dict.items():
        r = int(k) - lo + 1
        if 1 <= r <= n:
            out[r] = np.asarray(v, np.int32)[:width]
    return out


def build_2hop(hop1_off, hop1_ids, n_user, n_item, n_slices, start_time=0, max_1hop=10, max_2hop=100, seed=11, device=0):
    """GraphStore.construct_coll_2hop (code/graph_storage.py:127-246) on the GPU (csrc/hop2.cu).
    -> (hop1_ids with the reference's in-place shuffles applied, hop2_off, hop2_ids, hop2_deg): the CSR arrays
    ``GraphStore`` below takes.  No CPU fallback: raises RuntimeError without a device."""
    lib = _capi.load()
    off = np.ascontiguousarray(hop1_off, np.int64)
    ids = np.ascontiguousarray(hop1_ids, np.int32)
    n_lists = (n_user + n_item + 1) * n_slices
    if off.size != n_lists + 1:
        raise ValueError("hop1_off must have (n_user + n_item + 1) * n_slices + 1 entries")
    d = _capi.ScoreHop2Desc(int(n_user), int(n_item), int(n_slices), int(start_time), int(max_1hop), int(max_2hop),
                            off.ctypes.data, ids.ctypes.data if ids.size else None, int(seed))
    ids_out = np.empty_like(ids)
    off2 = np.empty(n_lists + 1, np.int64)
    n2 = C.c_int64()

    def check(rc):
        if rc:
            msg = lib.score_graph_build_2hop_error().decode()
            raise (ValueError if rc in (_capi.ERR_ARG, _capi.ERR_ID_RANGE) else RuntimeError)(msg)

    check(lib.score_graph_build_2hop(C.byref(d), int(device), None, off2.ctypes.data, None, None, 0, C.byref(n2)))
    ids2 = np.empty(max(n2.value, 1), np.int32)
    deg2 = np.empty(max(n2.value, 1), np.int32)
    check(lib.score_graph_build_2hop(C.byref(d), int(device), ids_out.ctypes.data if ids.size else None, off2.ctypes.data,
                                     ids2.ctypes.data, deg2.ctypes.data, ids2.size, C.byref(n2)))
    return ids_out, off2, ids2[:n2.value], deg2[:n2.value]


class GraphStore(object):
    """The interaction graph in device memory."""

    def __init__(self, n_user, n_item, n_slices, hop1_off, hop1_ids, hop2_off, hop2_ids, hop2_deg=None,
                 user_feat=None, item_feat=None, user_fnum=1, item_fnum=1, device=0):
        self._lib = _capi.load()
        self._g = C.c_void_p()
        self.n_user, self.n_item, self.n_slices = int(n_user), int(n_item), int(n_slices)
        self.user_fnum, self.item_fnum = int(user_fnum), int(item_fnum)
        keep = [np.ascontiguousarray(hop1_off, np.int64), np.ascontiguousarray(hop1_ids, np.int32),
                np.ascontiguousarray(hop2_off, np.int64), np.ascontiguousarray(hop2_ids, np.int32),
                None if hop2_deg is None else np.ascontiguousarray(hop2_deg, np.int32),
                None if user_feat is None else np.ascontiguousarray(user_feat, np.int32),
                None if item_feat is None else np.ascontiguousarray(item_feat, np.int32)]
        n_off = (self.n_user + self.n_item + 1) * self.n_slices + 1
        if keep[0].size != n_off or keep[2].size != n_off:
            raise ValueError("offset arrays must have (n_user + n_item + 1) * n_slices + 1 entries")
        if self.user_fnum > 1 and (keep[5] is None or keep[5].shape != (self.n_user + 1, self.user_fnum - 1)):
            raise ValueError("user_feat must be [(n_user + 1), user_fnum - 1]")
        if self.item_fnum > 1 and (keep[6] is None or keep[6].shape != (self.n_item + 1, self.item_fnum - 1)):
            raise ValueError("item_feat must be [(n_item + 1), item_fnum - 1]")
        ptr = [None if a is None else a.ctypes.data for a in keep]
        desc = _capi.ScoreGraphDesc(self.n_user, self.n_item, self.n_slices, self.user_fnum, self.item_fnum, *ptr)
        rc = self._lib.score_graph_create(C.byref(desc), int(device), C.byref(self._g))
        if rc != 0:
            msg = self._lib.score_graph_last_error(None)
            self._g = C.c_void_p()
            msg = msg.decode() if msg else "score_graph_create failed (%d)" % rc
            raise (ValueError if rc == _capi.ERR_ARG else RuntimeError)(msg)
        self.device = int(device)

    def _check(self, rc):
        if rc == 0:
            return
        msg = self._lib.score_graph_last_error(self._g)
        msg = msg.decode() if msg else "error %d" % rc
        raise (ValueError if rc in (_capi.ERR_ARG, _capi.ERR_ID_RANGE) else RuntimeError)(msg)

    def sample(self, uids, iids, group, start_time, pred_time, max_time_len, obj_per_time_slice, mode="rs", seed=0,
               draw_id=0, stream=None, as_numpy=False):
        """-> the loader's 8-tuple (graph_loader.py:383).  Default: CUDA torch tensors viewing the store's output
        buffers (valid until the next call); as_numpy=True copies them to the host."""
        u = np.ascontiguousarray(np.asarray(uids, np.int32))
        i = np.ascontiguousarray(np.asarray(iids, np.int32))
        B = int(i.size)
        if u.size != (B + group - 1) // group:
            raise ValueError("need one uid per group of %d target items" % group)
        out = _capi.ScoreBatch()
        self._check(self._lib.score_graph_sample(self._g, u.ctypes.data, i.ctypes.data, B, int(group), int(start_time),
                                                 int(pred_time), int(max_time_len), int(obj_per_time_slice),
                                                 MODES[mode], int(seed), int(draw_id), stream, C.byref(out)))
        T, K, fu, fi = int(max_time_len), int(obj_per_time_slice), self.user_fnum, self.item_fnum
        shapes = [(B, T, K, fi), (B, T, K, fu), (B, T, K, fu), (B, T, K, fi), (B, fu), (B, fi), (B,), (B,)]
        ptrs = [out.user_1hop, out.user_2hop, out.item_1hop, out.item_2hop, out.target_user, out.target_item,
                out.label, out.length]
        if as_numpy:
            self._check(self._lib.score_graph_sync(self._g, stream))
            res = []
            for p, sh in zip(ptrs, shapes):
                a = np.empty(sh, np.int32)
                if self._lib.score_copy_to_host(a.ctypes.data, p, a.nbytes) != 0:
                    raise RuntimeError("device to host copy failed")
                res.append(a)
            return tuple(res)
        import torch
        dev = torch.device("cuda", self.device)

        class _View:
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i4", "data": (int(ptr), False), "version": 2}
        if stream is None:
            self._check(self._lib.score_graph_sync(self._g, None))   # the model runs on its own stream
        return tuple(torch.as_tensor(_View(p, sh), device=dev) for p, sh in zip(ptrs, shapes))

    def close(self):
        if getattr(self, "_g", None) and self._g.value:
            self._lib.score_graph_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceGraphLoader(object):
    """Iterates a target file like GraphLoader (graph_loader.py:279-403): each line ``uid,pos_iid,neg_iid...``; the
    first 1 + neg_sample_num item ids of a line are that user's samples (:326-331).  Yields device-resident 8-tuples."""

    def __init__(self, store, batch_size, target_lines, start_time, pred_time, neg_sample_num, max_time_len,
                 obj_per_time_slice, mode="rs", seed=1111, stream=None):
        grp = 1 + neg_sample_num
        if batch_size % grp != 0:
            raise ValueError("batch size should be time of {}".format(grp))   # graph_loader.py:289-291
        self.store, self.grp, self.lines_per_batch = store, grp, batch_size // grp
        self.start_time, self.pred_time, self.T, self.K = start_time, pred_time, max_time_len, obj_per_time_slice
        self.mode, self.seed, self.stream = mode, seed, stream
        self.uids, self.iids = [], []
        for line in target_lines:
            f = line.strip().split(",")
            if len(f) < 1 + grp:
                continue
            self.uids.append(int(f[0]))
            self.iids.append([int(x) for x in f[1:1 + grp]])
        self.num_of_batch = (len(self.uids) + self.lines_per_batch - 1) // self.lines_per_batch
        self._next = 0

    def __iter__(self):
        self._next = 0
        return self

    def __next__(self):
        if self._next >= self.num_of_batch:
            raise StopIteration
        lo = self._next * self.lines_per_batch
        hi = min(lo + self.lines_per_batch, len(self.uids))
        uids = np.asarray(self.uids[lo:hi], np.int32)
        iids = np.asarray(self.iids[lo:hi], np.int32).reshape(-1)
        batch = self.store.sample(uids, iids, self.grp, self.start_time, self.pred_time, self.T, self.K, self.mode,
                                  seed=self.seed, draw_id=self._next, stream=self.stream)
        self._next += 1
        return batch


# ------------------------------------------------------------------------------------------ reference-signature loader
_GRAPHS = {}


def register_graph(db_name, store):
    """The reference addresses its graph by MongoDB database name (graph_handler_params[1], train_score.py:292-297:
    'tmall_2hop', 'taobao_2hop', 'ccmr_2hop'); here that name maps to a GraphStore in device memory."""
    _GRAPHS[str(db_name)] = store


class _HostReadable(object):
    """A device-resident id tensor that NumPy can read as well: train_score.py:157 takes the target item ids of a batch
    with np.array(batch_data[5])[:, 0].  Everything else (is_cuda, data_ptr, shape, ...) is the tensor's own."""

    def __init__(self, tensor):
        self._t = tensor

    def __array__(self, dtype=None, copy=None):
        a = self._t.cpu().numpy() if hasattr(self._t, "cpu") else np.asarray(self._t)
        return a.astype(dtype) if dtype is not None else a

    def __getattr__(self, name):
        return getattr(self._t, name)

    def __len__(self):
        return len(self._t)


class GraphLoader(DeviceGraphLoader):
    """GraphLoader of the reference by its own constructor signature (graph_loader.py:279-281):

        GraphLoader(graph_handler_params, batch_size, target_file, start_time, pred_time, worker_n, neg_sample_num)

    with graph_handler_params the list train_score.py builds (:292-297: time_slice_num, db_name, obj_per_time_slice,
    user_num, item_num, start_time, user_per_collection, item_per_collection, mode, user / item feature files, user_fnum,
    item_fnum).  db_name selects a GraphStore registered with register_graph(); the Mongo sharding parameters, the
    feature files (the store holds the tables) and worker_n (one kernel, no worker processes) are accepted and unused.
    Iterating yields the reference's 8-element batches with the id tensors in device memory."""

    _instances = 0      # every loader the caller builds (one per epoch / evaluation, train_score.py:151,214) draws anew

    def __init__(self, graph_handler_params, batch_size, target_file, start_time, pred_time, worker_n, neg_sample_num,
                 max_q_size=10, wait_time=0.01, seed=1111):
        p = list(graph_handler_params)
        if len(p) < 9:
            raise ValueError("graph_handler_params must be the list of train_score.py:292-297")
        time_slice_num, db_name, obj_per_time_slice, mode = int(p[0]), str(p[1]), int(p[2]), p[8]
        if db_name not in _GRAPHS:
            raise KeyError("no graph registered under %r: call score_b200.graph.register_graph(%r, store) first" % (db_name, db_name))
        if batch_size % (1 + neg_sample_num) != 0:
            print('batch size should be time of {}'.format(1 + neg_sample_num))        # graph_loader.py:289-291
            raise SystemExit(1)
        with open(target_file, 'r') as f:
            lines = f.readlines()
        DeviceGraphLoader.__init__(self, _GRAPHS[db_name], batch_size, lines, start_time, pred_time, neg_sample_num,
                                   time_slice_num - start_time - 1, obj_per_time_slice, mode=mode,
                                   seed=seed + GraphLoader._instances)
        GraphLoader._instances += 1
        self.batch_size, self.worker_n = batch_size, worker_n

    def __next__(self):
        batch = list(DeviceGraphLoader.__next__(self))
        batch[5] = _HostReadable(batch[5])
        return batch

    def stop(self):
        pass
