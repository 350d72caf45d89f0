"""Multi-GPU parity under pytest: the real-NCCL data-parallel and row-sharded steps against one GPU stepping on the
global batch (tools/multigpu_check.py under torchrun; skipped on a box with fewer GPUs), and the row-sharded trainer
at world size 1 (a real NCCL group of one rank) against the plain step."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

import parity_util as pu
from score_b200 import model as sb
from score_b200.synth import SHAPES, make_batch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(n, script, *args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, script)] + list(args)
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    sys.stdout.write(r.stdout[-6000:])
    sys.stderr.write(r.stderr[-3000:])
    return r


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("variant", ["plain", "graph_dropout"])
def test_real_nccl_dp_and_sharded_match_one_gpu(variant):
    """2 ranks (or every GPU of the box up to 8): losses, tables, slots and dense variables of the data-parallel and the
    row-sharded step vs a single GPU on the concatenated batch; replicas bit-identical.  `graph_dropout` runs the
    production configuration (CUDA-graph half-step, keep_prob 0.8, batch sizes that change between steps)."""
    n = min(_ngpu(), 8)
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = _torchrun(n, "tools/multigpu_check.py", "tiny_tb", variant)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIGPU_CHECK OK" in r.stdout


@pytest.fixture
def nccl_world1():
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    yield dist
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["lazy", "dense"])
def test_sharded_trainer_world1_equals_plain_step(nccl_world1, mode):
    """ShardedEmbeddingTrainer with one rank: owner = id % 1, local row = id + 1 - the same arithmetic as the plain step
    on a table shifted by one row; losses and every variable must agree (the exchange is an identity all-to-all)."""
    from score_b200 import parallel
    shape = SHAPES["tiny_tb"]
    lr, lam = 1e-3, 1e-4
    batches = [make_batch(shape, seed=900 + i) for i in range(4)]
    cfg, params, m1 = pu.make_models(shape, adam_mode=mode)
    l1 = [m1.train(None, b, lr, lam, keep_prob=1.0) for b in batches]
    a = list(shape.ctor_args())
    a[0] = parallel.shard_rows(shape.feature_size, 1)
    m2 = sb.SCORE(*a, adam_mode=mode, init_weights=False, use_graph=False)
    for name, _ in m2.tensor_names():
        v = params[name]
        if name == "emb_mtx":
            v = parallel.global_to_local_table(v, 1, 0)
        m2.set_tensor(name, v.numpy())
    sh = parallel.ShardedEmbeddingTrainer(m2, 1, 0)
    l2 = [sh.train(None, b, lr, lam, keep_prob=1.0) for b in batches]
    assert pu.rel_err(l2, l1) <= 1e-6
    for suf in ("", "/Adam", "/Adam_1"):
        got = m2.get_tensor("emb_mtx" + suf)[1:]
        want = m1.get_tensor("emb_mtx" + suf)
        assert np.array_equal(got[1:], want[1:]), "emb_mtx" + suf     # same rows, same order of additions: bit-exact
    for name in ("fc1/kernel", "dense_3/kernel", "gru_user_side/gru_cell/gates/kernel"):
        assert np.array_equal(m2.get_tensor(name), m1.get_tensor(name)), name
    p1, _, e1 = m1.eval(None, batches[0], lam)
    p2, _, e2 = sh.eval(None, batches[0], lam)
    assert pu.rel_err(p2, p1) <= 1e-6 and abs(e1 - e2) <= 1e-6 * abs(e1)
    m1.close(); m2.close()
