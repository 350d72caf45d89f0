"""ctypes binding of include/score_b200.h.  Loads the in-tree libscore_b200.so; there is no fallback:
a missing library is an ImportError-class failure at first use, a missing GPU a RuntimeError."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None


class ScoreConfig(C.Structure):
    _fields_ = [("feature_size", C.c_int64), ("eb_dim", C.c_int32), ("hidden_size", C.c_int32),
                ("max_time_len", C.c_int32), ("obj_per_time_slice", C.c_int32), ("user_fnum", C.c_int32),
                ("item_fnum", C.c_int32), ("model_type", C.c_int32), ("adam_mode", C.c_int32),
                ("max_batch", C.c_int32), ("seed", C.c_uint64), ("init_weights", C.c_int32),
                ("use_graph", C.c_int32)]


class ScoreBatch(C.Structure):
    _fields_ = [("user_1hop", C.c_void_p), ("user_2hop", C.c_void_p), ("item_1hop", C.c_void_p),
                ("item_2hop", C.c_void_p), ("target_user", C.c_void_p), ("target_item", C.c_void_p),
                ("label", C.c_void_p), ("length", C.c_void_p), ("batch_size", C.c_int32),
                ("on_device", C.c_int32)]


class ScoreGraphDesc(C.Structure):
    _fields_ = [("n_user", C.c_int32), ("n_item", C.c_int32), ("n_slices", C.c_int32), ("user_fnum", C.c_int32),
                ("item_fnum", C.c_int32), ("hop1_off", C.c_void_p), ("hop1_ids", C.c_void_p), ("hop2_off", C.c_void_p),
                ("hop2_ids", C.c_void_p), ("hop2_deg", C.c_void_p), ("user_feat", C.c_void_p), ("item_feat", C.c_void_p)]


class ScoreHop2Desc(C.Structure):
    _fields_ = [("n_user", C.c_int32), ("n_item", C.c_int32), ("n_slices", C.c_int32), ("start_time", C.c_int32),
                ("max_1hop", C.c_int32), ("max_2hop", C.c_int32), ("hop1_off", C.c_void_p), ("hop1_ids", C.c_void_p),
                ("seed", C.c_uint64)]


class ScoreShardPlan(C.Structure):
    _fields_ = [("counts", C.c_void_p), ("send_rows", C.c_void_p), ("staged", C.c_void_p), ("mini_keys", C.c_void_p),
                ("grad_send", C.c_void_p), ("n_positions", C.c_int64)]


# every symbol include/score_b200.h declares: (restype, argtypes)
_H = C.c_void_p
_F = C.c_float
SYMBOLS = {
    "score_create": (C.c_int, [C.POINTER(ScoreConfig), C.c_int, C.POINTER(_H)]),
    "score_destroy": (C.c_int, [_H]),
    "score_last_error": (C.c_char_p, [_H]),
    "score_train_step": (C.c_int, [_H, C.POINTER(ScoreBatch), _F, _F, _F, C.POINTER(_F)]),
    "score_train_step_async": (C.c_int, [_H, C.POINTER(ScoreBatch), _F, _F, _F]),
    "score_wait": (C.c_int, [_H, C.POINTER(_F)]),
    "score_eval": (C.c_int, [_H, C.POINTER(ScoreBatch), _F, C.c_void_p, C.POINTER(_F)]),
    "score_forward_backward": (C.c_int, [_H, C.POINTER(ScoreBatch), _F, _F, C.POINTER(_F)]),
    "score_get_buffer": (C.c_int, [_H, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "score_tensor_count": (C.c_int, [_H]),
    "score_tensor_info": (C.c_int, [_H, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "score_get_tensor": (C.c_int, [_H, C.c_char_p, C.c_void_p, C.c_size_t]),
    "score_set_tensor": (C.c_int, [_H, C.c_char_p, C.c_void_p, C.c_size_t]),
    "score_get_rows": (C.c_int, [_H, C.c_char_p, C.c_int64, C.c_int64, C.c_void_p]),
    "score_set_rows": (C.c_int, [_H, C.c_char_p, C.c_int64, C.c_int64, C.c_void_p]),
    "score_save": (C.c_int, [_H, C.c_char_p]),
    "score_restore": (C.c_int, [_H, C.c_char_p]),
    "score_eval_metrics": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "score_prepare_batch": (C.c_int, [_H, C.POINTER(ScoreBatch)]),
    "score_device_buffer": (C.c_int, [_H, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "score_gather_rows": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_void_p]),
    "score_step_begin": (C.c_int, [_H, C.POINTER(ScoreBatch), _F, _F, _F, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "score_dp_local_count": (C.c_int, [_H, C.POINTER(C.c_int32)]),
    "score_dp_block_words": (C.c_int64, [_H, C.c_int64]),
    "score_dp_pack": (C.c_int, [_H, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "score_dp_finish": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double)]),
    "score_dp_push": (C.c_int, [_H, C.c_int64, C.POINTER(C.c_uint64), C.c_int32, C.c_int64]),
    "score_step_finish": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "score_shard_plan": (C.c_int, [_H, C.c_int32, C.POINTER(ScoreShardPlan)]),
    "score_shard_pack_grads": (C.c_int, [_H]),
    "score_shard_counts_fetch": (C.c_int, [_H, C.c_void_p, C.c_int32]),
    "score_shard_counts_wait": (C.c_int, [_H, C.c_void_p, C.c_int32]),
    "score_shard_presort": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "score_shard_serve_push": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]),
    "score_shard_grad_push": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]),
    "score_shard_register_staged": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "score_set_sample_offset": (C.c_int, [_H, C.c_int32]),
    "score_stream": (C.c_int, [_H, C.POINTER(C.c_void_p)]),
    "score_launch_count": (C.c_int64, [_H]),
    "score_enable_probes": (C.c_int, [_H, C.c_int]),
    "score_probe_times": (C.c_int, [_H, C.c_void_p, C.c_int]),
    "score_last_step_stats": (C.c_int, [_H, C.c_void_p]),
    "score_graph_create": (C.c_int, [C.POINTER(ScoreGraphDesc), C.c_int, C.POINTER(_H)]),
    "score_graph_destroy": (C.c_int, [_H]),
    "score_graph_last_error": (C.c_char_p, [_H]),
    "score_graph_sample": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_uint64, C.c_uint32, C.c_void_p, C.POINTER(ScoreBatch)]),
    "score_graph_sync": (C.c_int, [_H, C.c_void_p]),
    "score_graph_build_2hop": (C.c_int, [C.POINTER(ScoreHop2Desc), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int64, C.POINTER(C.c_int64)]),
    "score_graph_build_2hop_error": (C.c_char_p, []),
    "score_copy_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
}

ERR_ARG, ERR_CUDA, ERR_ID_RANGE, ERR_IO, ERR_NAME = 1, 2, 3, 4, 5


def lib_path() -> str:
    return _build.LIB


def load():
    """dlopen the library (building it first when nvcc is present and it is missing/stale)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if not os.path.exists(path):
        _build.build_library()
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(lib, handle, rc):
    if rc == 0:
        return
    msg = lib.score_last_error(handle)
    msg = msg.decode() if msg else "error %d" % rc
    if rc in (ERR_ARG, ERR_ID_RANGE, ERR_NAME):
        raise ValueError(msg)
    if rc == ERR_IO:
        raise IOError(msg)
    raise RuntimeError(msg)
