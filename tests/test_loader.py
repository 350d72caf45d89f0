"""The on-GPU graph store + neighbor sampler (csrc/sampler.cu) and its CPU restatement
(oracle/loader_ref.py) against tests/golden/loader_reference.npz - batches produced by the reference's OWN
GraphHandler code (graph_loader.py:40-277, executed unmodified by tools/make_golden.py).  ids are integers: bit-exact."""
import os

import numpy as np
import pytest

from oracle import loader_ref as L
from score_b200.synth import SHAPES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loader_reference.npz")


def _cases():
    return [str(c) for c in np.load(GOLDEN)["cases"]]


def _case(g, name):
    nu, ni, tsn, start, pred, K, uf, fi, mode, neg, seed, draw = (int(x) for x in g[name + "/params"])
    return dict(nu=nu, ni=ni, tsn=tsn, start=start, pred=pred, K=K, uf=uf, fi=fi, mode="is" if mode else "rs", neg=neg,
                seed=seed, draw=draw, T=tsn - start - 1)


def _docs_from_csr(g, name, c):
    """rebuild the per-node documents of the reference from the stored CSR arrays"""
    off1, ids1 = g[name + "/hop1_off"], g[name + "/hop1_ids"]
    off2, ids2, deg2 = g[name + "/hop2_off"], g[name + "/hop2_ids"], g[name + "/hop2_deg"]
    S = c["tsn"]
    user_docs, item_docs = {}, {}
    for node in range(1, c["nu"] + c["ni"] + 1):
        d = {"1hop": [], "2hop": [], "degrees": []}
        for s in range(S):
            r = node * S + s
            d["1hop"].append(ids1[off1[r]:off1[r + 1]].tolist())
            d["2hop"].append(ids2[off2[r]:off2[r + 1]].tolist())
            d["degrees"].append(deg2[off2[r]:off2[r + 1]].tolist())
        (user_docs if node <= c["nu"] else item_docs)[node] = d
    uft, ift = g[name + "/user_feat"], g[name + "/item_feat"]
    ufd = None if c["uf"] == 1 else {str(u): uft[u].tolist() for u in range(1, c["nu"] + 1)}
    ifd = None if c["fi"] == 1 else {str(c["nu"] + r): ift[r].tolist() for r in range(1, c["ni"] + 1)}
    return user_docs, item_docs, ufd, ifd


@pytest.mark.parametrize("name", _cases())
def test_loader_oracle_matches_the_references_own_graph_handler(name):
    g = np.load(GOLDEN)
    c = _case(g, name)
    user_docs, item_docs, ufd, ifd = _docs_from_csr(g, name, c)
    uni = lambda side, ent, ts: L.draw_uniforms(c["seed"], c["draw"], side, ent, ts, c["T"], c["K"])
    gh = L.GraphHandlerRef(c["tsn"], user_docs, item_docs, c["K"], c["nu"], c["ni"], c["start"], c["mode"], ufd, ifd,
                           c["uf"], c["fi"], uni)
    batch = L.assemble_batch(gh, g[name + "/uids"].tolist(), g[name + "/iids"].tolist(), c["pred"], c["start"], c["neg"])
    for k in range(8):
        got = np.asarray(batch[k]).astype(np.int32)
        assert np.array_equal(got, g[name + "/batch/%d" % k]), "tensor %d of %s" % (k, name)


def test_philox_restatement_known_answer():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors: counter 0 / key 0 and the all-ones case)"""
    def block(ctr, key):
        c = [np.array([x], np.uint64) for x in ctr]
        k0, k1 = np.uint64(key[0]), np.uint64(key[1])
        for _ in range(10):
            p0, p1 = np.uint64(L.M0) * c[0], np.uint64(L.M1) * c[2]
            m = np.uint64(L.MASK)
            c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & m, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & m]
            k0, k1 = (k0 + np.uint64(L.W0)) & m, (k1 + np.uint64(L.W1)) & m
        return [int(x[0]) for x in c]
    assert block((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert block((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    u = L.philox_uniform(1111, 1, 0, np.arange(64))
    assert u.dtype == np.float32 and (u >= 0).all() and (u < 1).all()


def test_graph_store_needs_a_gpu_and_validates_arguments():
    import torch
    from score_b200.graph import GraphStore
    off = np.zeros(3 * 2 + 1, np.int64)
    with pytest.raises(ValueError):
        GraphStore(1, 1, 2, off[:-1], [], off, [])                      # offset array too short
    with pytest.raises(ValueError):
        GraphStore(1, 1, 2, off, [], off, [], user_fnum=3)              # missing user feature table
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            GraphStore(1, 1, 2, off, [], off, [])                       # no CPU fallback


def test_graph_store_checks_the_neighbor_types_on_the_host():
    """score_graph_create validates the CSR arrays before it touches the device: the sampler indexes the side-feature
    tables with the neighbor ids (graph_loader.py:186-191 raises KeyError for an id without an entry).  Every graph the
    GPU tests use passes the check (on a CPU box construction then stops at 'no CUDA device'); a mistyped id does not."""
    import torch
    from score_b200.graph import GraphStore, docs_to_csr, feat_table
    if torch.cuda.is_available():
        pytest.skip("CPU-side check of the validation order")
    g = np.load(GOLDEN)
    for name in _cases():
        with pytest.raises(RuntimeError, match="no CUDA device"):
            _store(g, name, _case(g, name))
    for seed, (nu, ni, tsn, uf, fi, kw) in {9: (400, 600, 12, 3, 4, {}), 12: (300, 500, 9, 1, 2, {"n_feat": 40})}.items():
        user_docs, item_docs, ufd, ifd = L.random_graph(np.random.default_rng(seed), nu, ni, tsn, user_fnum=uf, item_fnum=fi, **kw)
        off1, ids1, off2, ids2, deg2 = docs_to_csr(user_docs, item_docs, nu, ni, tsn)
        args = (nu, ni, tsn, off1, ids1, off2, ids2, deg2, feat_table(ufd, 1, nu, uf - 1) if uf > 1 else None,
                feat_table(ifd, nu + 1, ni, fi - 1) if fi > 1 else None, uf, fi)
        with pytest.raises(RuntimeError, match="no CUDA device"):
            GraphStore(*args)
        u_first = int(off1[1 * tsn])                       # first 1-hop entry of user 1: must be an item id
        if int(off1[(nu + 1) * tsn]) > u_first:
            bad = ids1.copy(); bad[u_first] = 1            # a user id in a user's 1-hop list
            with pytest.raises(ValueError, match="hop1 list of node"):
                GraphStore(*(args[:4] + (bad,) + args[5:]))
        if ids2.size:
            bad2 = ids2.copy(); bad2[0] = nu + ni + 1      # outside the node range
            with pytest.raises(ValueError, match="hop2 list of node"):
                GraphStore(*(args[:6] + (bad2,) + args[7:]))
        dec = off1.copy(); dec[5] = dec[-1] + 1
        with pytest.raises(ValueError, match="ascending"):
            GraphStore(*(args[:3] + (dec,) + args[4:]))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import run_tmall_sample as rt
    gt, nu, ni, V, S = rt.load_fixture()
    with pytest.raises(RuntimeError, match="no CUDA device"):
        GraphStore(nu, ni, S, gt["hop1_off"], gt["hop1_ids"], gt["hop2_off"], gt["hop2_ids"], gt["hop2_deg"], gt["user_feat"],
                   gt["item_feat"], rt.UF, rt.IF)


# ------------------------------------------------------------------------------------------ GPU
def _store(g, name, c):
    from score_b200.graph import GraphStore
    return GraphStore(c["nu"], c["ni"], c["tsn"], g[name + "/hop1_off"], g[name + "/hop1_ids"], g[name + "/hop2_off"],
                      g[name + "/hop2_ids"], g[name + "/hop2_deg"],
                      g[name + "/user_feat"] if c["uf"] > 1 else None, g[name + "/item_feat"] if c["fi"] > 1 else None,
                      c["uf"], c["fi"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", _cases())
def test_cuda_sampler_matches_the_references_own_graph_handler(name):
    g = np.load(GOLDEN)
    c = _case(g, name)
    st = _store(g, name, c)
    got = st.sample(g[name + "/uids"], g[name + "/iids"], c["neg"] + 1, c["start"], c["pred"], c["T"], c["K"], c["mode"],
                    seed=c["seed"], draw_id=c["draw"], as_numpy=True)
    for k in range(8):
        assert np.array_equal(got[k], g[name + "/batch/%d" % k]), "tensor %d of %s" % (k, name)
    st.close()


@pytest.mark.gpu
def test_cuda_sampler_matches_oracle_on_a_larger_graph_and_rejects_unknown_targets():
    from score_b200.graph import GraphStore, docs_to_csr, feat_table
    rng = np.random.default_rng(9)
    nu, ni, tsn, K, uf, fi, neg = 400, 600, 12, 10, 3, 4, 1
    user_docs, item_docs, ufd, ifd = L.random_graph(rng, nu, ni, tsn, user_fnum=uf, item_fnum=fi)
    off1, ids1, off2, ids2, deg2 = docs_to_csr(user_docs, item_docs, nu, ni, tsn)
    st = GraphStore(nu, ni, tsn, off1, ids1, off2, ids2, deg2, feat_table(ufd, 1, nu, uf - 1),
                    feat_table(ifd, nu + 1, ni, fi - 1), uf, fi)
    start, pred, T = 0, 9, tsn - 1
    uids = rng.integers(1, nu + 1, 128)
    iids = rng.integers(nu + 1, nu + ni + 1, 256)
    for mode in ("rs", "is"):
        got = st.sample(uids, iids, neg + 1, start, pred, T, K, mode, seed=5, draw_id=3, as_numpy=True)
        uni = lambda side, ent, ts: L.draw_uniforms(5, 3, side, ent, ts, T, K)
        gh = L.GraphHandlerRef(tsn, user_docs, item_docs, K, nu, ni, start, mode, ufd, ifd, uf, fi, uni)
        want = L.assemble_batch(gh, uids.tolist(), iids.tolist(), pred, start, neg)
        for k in range(8):
            assert np.array_equal(got[k], np.asarray(want[k]).astype(np.int32)), (mode, k)
    bad = iids.copy()
    bad[7] = 5   # a user id where an item id belongs
    with pytest.raises(ValueError):
        st.sample(uids, bad, neg + 1, start, pred, T, K, as_numpy=True)
    st.close()


@pytest.mark.gpu
def test_device_loader_feeds_the_model_without_host_copies():
    """DeviceGraphLoader -> SCORE.train / eval with device-resident ids gives the same numbers as the same ids fed
    from the host (the boundary accepts both, model.py:_Batch)."""
    from score_b200 import model as sb
    from score_b200.graph import DeviceGraphLoader, GraphStore, docs_to_csr
    rng = np.random.default_rng(12)
    shape = SHAPES["tiny_tb"]          # uf/if = 1/2, T=8, K=10
    nu, ni, tsn = 300, 500, shape.max_time_len + 1
    user_docs, item_docs, _, ifd = L.random_graph(rng, nu, ni, tsn, user_fnum=1, item_fnum=2, n_feat=40)
    off1, ids1, off2, ids2, deg2 = docs_to_csr(user_docs, item_docs, nu, ni, tsn)
    from score_b200.graph import feat_table
    st = GraphStore(nu, ni, tsn, off1, ids1, off2, ids2, deg2, None, feat_table(ifd, nu + 1, ni, 1), 1, 2)
    lines = ["%d,%s" % (rng.integers(1, nu + 1), ",".join(str(x) for x in rng.integers(nu + 1, nu + ni + 1, 2)))
             for _ in range(40)]
    V = nu + ni + 1 + 40
    args = (V,) + shape.ctor_args()[1:]
    m_dev = sb.SCORE(*args, seed=3, use_graph=False)
    m_host = sb.SCORE(*args, seed=3, use_graph=False)
    loader = DeviceGraphLoader(st, 32, lines, 0, 6, 1, shape.max_time_len, shape.obj_per_time_slice, seed=21)
    n = 0
    for batch in loader:
        host = tuple(x.cpu().numpy() for x in batch)
        l_dev = m_dev.train(None, batch, 1e-3, 1e-4, keep_prob=1.0)
        l_host = m_host.train(None, host, 1e-3, 1e-4, keep_prob=1.0)
        assert l_dev == l_host
        n += 1
    assert n == loader.num_of_batch == 3      # 40 lines, 16 per batch: the last batch is short
    st.close(); m_dev.close(); m_host.close()
