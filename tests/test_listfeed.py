"""The nested-list feed of the Python mirror (score_b200/csrc/listfeed.c): same int32 arrays as NumPy's converter on the
loader's output format (graph_loader.py:383: nested lists, float zeros in dummy slices), same errors on malformed input."""
import numpy as np
import pytest

from score_b200 import model as sb
from score_b200.synth import SHAPES, make_batch

lf = sb._listfeed
pytestmark = pytest.mark.skipif(lf is None, reason="no C compiler: the NumPy converter is in use")


def cfg_of(batch):
    return dict(max_time_len=batch[0].shape[1], obj_per_time_slice=batch[0].shape[2], user_fnum=batch[1].shape[3],
                item_fnum=batch[0].shape[3])


@pytest.mark.parametrize("shape", ["tiny", "tiny_tb", "tmall"])
def test_lists_convert_like_numpy(shape):
    b = make_batch(SHAPES[shape], seed=4)
    lists = [x.tolist() for x in b]
    # the loader's dummy slices are float zeros (np.zeros(...).tolist(), graph_loader.py:90-91)
    for s in range(0, len(lists[0]), 3):
        lists[0][s][0] = np.zeros(b[0].shape[2:]).tolist()
        lists[3][s][-1] = np.zeros(b[3].shape[2:]).tolist()
    want = [np.asarray(x).astype(np.int32) for x in lists]
    got = sb._Batch(tuple(lists), cfg_of(b))
    for w, g in zip(want, got.keep):
        assert g.dtype == np.int32 and g.flags.c_contiguous and np.array_equal(w, g)
    assert got.B == b[0].shape[0]


def test_leaf_types_and_truncation():
    out = np.empty((2, 4), np.int32)
    lf.fill_i32([[1, 2.9, -2.9, True], (np.int64(7), np.float32(3.5), np.int32(-4), 0.0)], (2, 4), out)
    assert out.tolist() == [[1, 2, -2, 1], [7, 3, -4, 0]]          # truncation toward zero, like ndarray.astype(int32)
    e = np.empty((0, 3), np.int32)
    lf.fill_i32([], (0, 3), e)


def test_malformed_input_raises():
    out = np.empty((2, 2), np.int32)
    with pytest.raises(ValueError):
        lf.fill_i32([[1, 2], [3]], (2, 2), out)                     # ragged
    with pytest.raises(ValueError):
        lf.fill_i32([[1, 2], [3, [4]]], (2, 2), out)                # nested deeper than the shape
    with pytest.raises(ValueError):
        lf.fill_i32([1, 2], (2, 2), out)                            # not nested enough
    with pytest.raises(OverflowError):
        lf.fill_i32([[1, 2], [3, 2 ** 31]], (2, 2), out)
    with pytest.raises(ValueError):
        lf.fill_i32([[1, 2], [3, 4]], (2, 2), np.empty(3, np.int32))  # buffer size
    with pytest.raises((TypeError, ValueError, BufferError)):
        lf.fill_i32([[1, 2], [3, 4]], (2, 2), b"read-only-bytes!")   # not writable
    with pytest.raises(TypeError):
        lf.fill_i32([["a", 2], [3, 4]], (2, 2), out)


def test_batch_falls_back_to_numpy_for_lists_of_arrays_and_reports_bad_shapes():
    b = make_batch(SHAPES["tiny"], seed=1)
    mixed = tuple([np.asarray(r) for r in x] if i == 0 else x.tolist() for i, x in enumerate(b))   # list of per-sample arrays
    got = sb._Batch(mixed, cfg_of(b))
    assert np.array_equal(got.keep[0], b[0])
    bad = [x.tolist() for x in b]
    bad[4] = bad[4][:-1]                                             # one target user short
    with pytest.raises(ValueError):
        sb._Batch(tuple(bad), cfg_of(b))
