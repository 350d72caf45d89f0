"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/score_b200.h declares, the host mirror validates shapes, and nothing computes without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

from score_b200 import _capi
from score_b200 import model as sb
from score_b200.synth import SHAPES, make_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "score_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(score_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libscore_b200.so does not export %s" % name
    # and the ctypes table covers exactly the header
    assert sorted(_capi.SYMBOLS) == declared


def test_every_entry_point_is_documented_and_cites_the_reference():
    """INTEGRATION.md names every exported entry point (the table of what each one replaces in the reference), and the
    header cites reference file:line for the interface it stands in for."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in _declared_symbols() if n not in doc]
    assert not missing, missing
    header = open(os.path.join(ROOT, "include", "score_b200.h")).read()
    assert len(re.findall(r"[a-z_]+\.py:\d+", header)) >= 10


def test_struct_layout_matches_header(tmp_path):
    """Compile the real header with gcc and compare sizeof/offsetof of every struct with its ctypes mirror."""
    import ctypes as C
    import subprocess
    structs = ["ScoreConfig", "ScoreBatch", "ScoreGraphDesc", "ScoreHop2Desc", "ScoreShardPlan"]
    src = tmp_path / "layout.c"
    body = ""
    for name in structs:
        body += 'printf("%%zu\\n", sizeof(%s));' % name
        for f, _ in getattr(_capi, name)._fields_:
            body += 'printf("%%zu\\n", offsetof(%s, %s));' % (name, f)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "score_b200.h"\nint main(){' + body + 'return 0;}')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = []
    for name in structs:
        cls = getattr(_capi, name)
        want += [C.sizeof(cls)] + [getattr(cls, f).offset for f, _ in cls._fields_]
    assert out == want


def test_argument_validation_without_gpu():
    with pytest.raises(ValueError):
        sb.SCORE(100, 12, 32, 4, 10, 1, 2)        # eb_dim must be a power of two
    with pytest.raises(ValueError):
        sb.SCORE(100, 16, 32, 4, 40, 1, 2)        # K > 32
    with pytest.raises(ValueError):
        sb.SCORE(1, 16, 32, 4, 10, 1, 2)          # table needs the dummy row and one real row


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sb.SCORE(100, 16, 32, 4, 10, 1, 2)


def test_batch_marshalling_casts_and_checks_shapes():
    sh = SHAPES["tiny"]
    cfg = dict(max_time_len=sh.max_time_len, obj_per_time_slice=sh.obj_per_time_slice,
               user_fnum=sh.user_fnum, item_fnum=sh.item_fnum)
    b = make_batch(sh, seed=1, batch=6)
    mb = sb._Batch(b, cfg)
    assert mb.B == 6 and mb.struct.on_device == 0
    nested = [x.tolist() for x in b]
    nested[1][0][0] = np.zeros([sh.obj_per_time_slice, sh.user_fnum]).tolist()   # float dummy rows of the loader
    mb = sb._Batch(nested, cfg)
    assert all(k.dtype == np.int32 for k in mb.keep)
    bad = list(b)
    bad[0] = bad[0][:, :-1]
    with pytest.raises(ValueError):
        sb._Batch(bad, cfg)
    with pytest.raises(ValueError):
        sb._Batch(b[:7], cfg)


def test_synthetic_batches_follow_the_loader_contract():
    sh = SHAPES["tiny_tb"]
    b = make_batch(sh, seed=3)
    u1, u2, i1, i2, tu, ti, label, length = b
    B, T, K = sh.batch, sh.max_time_len, sh.obj_per_time_slice
    assert u1.shape == (B, T, K, sh.item_fnum) and u2.shape == (B, T, K, sh.user_fnum)
    assert i1.shape == (B, T, K, sh.user_fnum) and i2.shape == (B, T, K, sh.item_fnum)
    assert tu.shape == (B, sh.user_fnum) and ti.shape == (B, sh.item_fnum)
    assert label.tolist() == [1, 0] * (B // 2)                     # graph_loader.py:378-381
    assert np.array_equal(u1[0], u1[1]) and np.array_equal(tu[0], tu[1])   # user side repeated per (1+neg)
    assert (length == sh.length).all()
    assert np.array_equal(u1[:, sh.length:], np.repeat(u1[:, sh.length - 1:sh.length], T - sh.length, 1))
    for x in b[:6]:
        assert x.min() >= 0 and x.max() < sh.feature_size
    assert sh.ids_per_sample == 483 - 0 if sh.name == "taobao" else True
    assert SHAPES["taobao"].ids_per_sample == 483 and SHAPES["tmall"].ids_per_sample == 1547
    assert SHAPES["ccmr"].ids_per_sample == 4806 and SHAPES["ccmr_k20"].ids_per_sample == 9606
    assert SHAPES["large_vocab"].ids_per_sample == 322
