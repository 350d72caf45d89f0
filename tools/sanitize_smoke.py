"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): every model type fwd+bwd, train steps in the three
optimizer modes, eval, the sharded plan kernels, the sampler and the 2-hop builder."""
import ctypes as C
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from score_b200 import model as sb
from score_b200 import _capi
from score_b200.graph import build_2hop
from score_b200.synth import SHAPES, make_batch
for mode in ("dense", "lazy", "sparse"):
    for name in ("tiny", "tiny_tb"):      # tiny_tb takes the lean co-attention kernels (compiled-in Taobao geometry)
        sh = SHAPES[name]
        m = sb.SCORE(*sh.ctor_args(), adam_mode=mode, use_graph=False)
        b = make_batch(sh, seed=1)
        print(mode, name, "fwdbwd loss", m.forward_backward(b, 1e-4))
        for i in range(3):
            print(mode, name, "train loss", m.train(None, make_batch(sh, seed=2 + i), 5e-4, 1e-4))
        p, l, loss = m.eval(None, b, 1e-4)
        print(mode, name, "eval loss", loss, p[:3])
        if mode == "lazy":
            bb = sb._Batch(b, m.cfg)
            m._check(m._lib.score_prepare_batch(m._h, C.byref(bb.struct)))
            plan = _capi.ScoreShardPlan()
            m._check(m._lib.score_shard_plan(m._h, 3, C.byref(plan)))
            m._check(m._lib.score_shard_pack_grads(m._h))
        m.close()
sh = SHAPES["tiny"]
for mt in ("RIA", "RCA", "SCORE_USER", "SCORE_ITEM", "RRN"):
    m = getattr(sb, mt)(*sh.ctor_args(), use_graph=False)
    print(mt, "train loss", m.train(None, make_batch(sh, seed=9), 5e-4, 1e-4))
    m.close()
rng = np.random.default_rng(0)
nu, ni, S = 20, 16, 3
lens = rng.integers(0, 9, (nu + ni + 1) * S); lens[:S] = 0
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
ids = np.empty(off[-1], np.int32)
for n in range(1, nu + ni + 1):
    for t in range(S):
        a, b = off[n * S + t], off[n * S + t + 1]
        ids[a:b] = rng.integers(nu + 1, nu + ni + 1, b - a) if n <= nu else rng.integers(1, nu + 1, b - a)
out = build_2hop(off, ids, nu, ni, S, 0, 4, 9, 3)
print("hop2", len(out[2]))
print("SANITIZE_SMOKE_OK")
