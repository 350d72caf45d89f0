"""2-hop graph construction (code/graph_storage.py:127-246).

tests/golden/hop2_reference.npz holds documents produced by the reference's OWN GraphStore.construct_coll_2hop
(executed unmodified by tools/make_golden.py over in-memory collections, its two random calls fed the Philox
permutations).  CPU: the oracle restatement reproduces them.  GPU: the CUDA builder (csrc/hop2.cu) reproduces them,
and matches the oracle on a larger graph with long (multi-chunk) lists."""
import os

import numpy as np
import pytest

from oracle import graph_ref as G
from score_b200.graph import docs_to_csr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hop2_reference.npz")


def _csr_to_lists(off, ids, n_user, n_item, S):
    u = {n: [ids[off[n * S + t]:off[n * S + t + 1]].tolist() for t in range(S)] for n in range(1, n_user + 1)}
    i = {n: [ids[off[n * S + t]:off[n * S + t + 1]].tolist() for t in range(S)] for n in range(n_user + 1, n_user + n_item + 1)}
    return u, i


def _cases():
    g = np.load(GOLDEN, allow_pickle=False)
    return g, [str(c) for c in g["cases"]]


def _oracle_csr(off1, ids1, nu, ni, S, start, m1, m2, seed):
    u1, i1 = _csr_to_lists(off1, ids1, nu, ni, S)
    ud, idocs = G.build_2hop(u1, i1, nu, ni, S, start, m1, m2, seed)
    return docs_to_csr(ud, idocs, nu, ni, S)


def test_fixture_is_reference_produced():
    g, cases = _cases()
    assert str(g["source"]).startswith("reference:") and len(cases) >= 3
    # the second and third case exercise both caps; the first uses the reference's constants (10 / 100)
    assert any(int(g[c + "/params"][5]) < int(g[c + "/params"][4]) ** 2 for c in cases)


@pytest.mark.parametrize("case", ["caps_10_100", "caps_4_9_start1", "caps_3_5"])
def test_oracle_matches_the_references_own_construction(case):
    g, _ = _cases()
    nu, ni, S, start, m1, m2, seed = (int(x) for x in g[case + "/params"])
    off1, ids1_out, off2, ids2, deg2 = _oracle_csr(g[case + "/hop1_off"], g[case + "/hop1_ids"], nu, ni, S, start, m1, m2, seed)
    assert np.array_equal(ids1_out, g[case + "/hop1_ids_out"])      # the in-place shuffles of the long lists
    assert np.array_equal(off2, g[case + "/hop2_off"])
    assert np.array_equal(ids2, g[case + "/hop2_ids"])
    assert np.array_equal(deg2, g[case + "/hop2_deg"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["caps_10_100", "caps_4_9_start1", "caps_3_5"])
def test_cuda_builder_matches_the_references_own_construction(case):
    from score_b200.graph import build_2hop
    g, _ = _cases()
    nu, ni, S, start, m1, m2, seed = (int(x) for x in g[case + "/params"])
    ids1_out, off2, ids2, deg2 = build_2hop(g[case + "/hop1_off"], g[case + "/hop1_ids"], nu, ni, S, start, m1, m2, seed)
    assert np.array_equal(ids1_out, g[case + "/hop1_ids_out"])
    assert np.array_equal(off2, g[case + "/hop2_off"])
    assert np.array_equal(ids2, g[case + "/hop2_ids"])
    assert np.array_equal(deg2, g[case + "/hop2_deg"])


@pytest.mark.gpu
def test_cuda_builder_matches_the_oracle_on_long_lists():
    """lists of several thousand entries (more than one 2048-entry chunk of the rank kernel), empty nodes, start_time 2"""
    from score_b200.graph import build_2hop
    rng = np.random.default_rng(5)
    nu, ni, S, start, m1, m2, seed = 300, 200, 4, 2, 10, 60, 99
    u1, i1 = G.random_1hop(rng, nu, ni, S, 60000, hot_items=2, hot_share=0.5)
    in_u = {u: {"1hop": u1[u], "2hop": [[] for _ in range(S)], "degrees": [[] for _ in range(S)]} for u in u1}
    in_i = {i: {"1hop": i1[i], "2hop": [[] for _ in range(S)], "degrees": [[] for _ in range(S)]} for i in i1}
    off1, ids1, _, _, _ = docs_to_csr(in_u, in_i, nu, ni, S)
    assert (np.diff(off1) > 2048).any()
    want = _oracle_csr(off1, ids1, nu, ni, S, start, m1, m2, seed)
    ids1_out, off2, ids2, deg2 = build_2hop(off1, ids1, nu, ni, S, start, m1, m2, seed)
    assert np.array_equal(ids1_out, want[1])
    assert np.array_equal(off2, want[2]) and np.array_equal(ids2, want[3]) and np.array_equal(deg2, want[4])


@pytest.mark.gpu
def test_cuda_builder_rejects_bad_arguments():
    from score_b200.graph import build_2hop
    with pytest.raises(ValueError):
        build_2hop(np.zeros(3, np.int64), np.zeros(0, np.int32), 1, 1, 2)                  # hop1_off has the wrong size
    with pytest.raises(ValueError):
        build_2hop(np.zeros(5, np.int64), np.zeros(0, np.int32), 1, 1, 2, max_1hop=64)     # max_1hop > 32


def test_builder_validates_the_csr_arrays_on_the_host():
    """argument and CSR validation happens before any CUDA call (no GPU needed): sizes, caps, ascending offsets, neighbor
    ids inside 0..n_user+n_item (the kernels index the offset array with them; the reference raises IndexError on a
    neighbor without a document, graph_storage.py:160,208)"""
    from score_b200.graph import build_2hop
    nu, ni, S = 2, 2, 2
    off = np.array([0, 0, 0, 1, 2, 2, 3, 4, 4, 5, 5], np.int64)          # (nu + ni + 1) * S + 1 entries
    ids = np.array([3, 4, 3, 1, 2], np.int32)
    with pytest.raises(ValueError):
        build_2hop(off[:-1], ids, nu, ni, S)                                # wrong number of offsets
    with pytest.raises(ValueError):
        build_2hop(off, ids, nu, ni, S, max_1hop=64)                        # max_1hop > 32
    bad = ids.copy(); bad[1] = nu + ni + 1
    with pytest.raises(ValueError, match="outside"):
        build_2hop(off, bad, nu, ni, S)
    bad = ids.copy(); bad[0] = -1
    with pytest.raises(ValueError, match="outside"):
        build_2hop(off, bad, nu, ni, S)
    dec = off.copy(); dec[4] = 0
    with pytest.raises(ValueError, match="ascending"):
        build_2hop(dec, ids, nu, ni, S)
    shifted = off.copy(); shifted[0] = 1
    with pytest.raises(ValueError, match="must be 0"):
        build_2hop(shifted, ids, nu, ni, S)
