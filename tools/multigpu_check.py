"""Run under torchrun on N GPUs: data-parallel and row-sharded steps vs a single-GPU run on the global batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import score_ref as ref          # weights only (the checker's initialiser)
from score_b200 import model as sb
from score_b200 import parallel
from score_b200.synth import SHAPES, make_batch


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shape = SHAPES[sys.argv[1] if len(sys.argv) > 1 else "tiny_tb"]
    # "plain": direct launches, keep_prob 1, one batch size.  "graph_dropout": the production configuration - CUDA-graph
    # half-step, dropout (masks keyed by the GLOBAL sample index, so they equal the single-GPU run's), batch sizes that
    # change from step to step (graph per size, exchange buffers regrow)
    variant = sys.argv[2] if len(sys.argv) > 2 else "plain"
    prod = variant == "graph_dropout"
    lr, lam = 1e-3, 1e-4
    kp = 0.8 if prod else 1.0
    sizes = [shape.batch, shape.batch // 2, shape.batch, shape.batch, shape.batch // 4, shape.batch] if prod else [shape.batch] * 3
    steps = len(sizes)
    cfg = ref.ScoreConfig(*shape.ctor_args())
    params = ref.init_params(cfg, 5)
    per = [[make_batch(shape, seed=100 * s + r, batch=sizes[s]) for r in range(world)] for s in range(steps)]
    glob = [tuple(np.concatenate([b[i] for b in bs], 0) for i in range(8)) for bs in per]
    worst = 0.0
    for mode in ("dense", "lazy"):
        # single-GPU run on the global batch
        m1 = sb.SCORE(*shape.ctor_args(), device=local, adam_mode=mode, init_weights=False, use_graph=False)
        m1.load_params(params)
        l1 = [m1.train(None, g, lr, lam, keep_prob=kp) for g in glob]
        # data-parallel, replicated table
        m2 = sb.SCORE(*shape.ctor_args(), device=local, adam_mode=mode, init_weights=False, use_graph=prod)
        m2.load_params(params)
        dp = parallel.DataParallelTrainer(m2, world, rank)
        l2 = [dp.train(None, bs[rank], lr, lam, keep_prob=kp) for bs in per]
        # row-sharded table
        a = list(shape.ctor_args())
        a[0] = parallel.shard_rows(shape.feature_size, world)
        m3 = sb.SCORE(*a, device=local, adam_mode=mode, init_weights=False, use_graph=prod)
        for name, _ in m3.tensor_names():
            v = params[name]
            if name == "emb_mtx":
                v = parallel.global_to_local_table(v, world, rank)
            m3.set_tensor(name, v.numpy())
        sh = parallel.ShardedEmbeddingTrainer(m3, world, rank)
        l3 = [sh.train(None, bs[rank], lr, lam, keep_prob=kp) for bs in per]
        p1, _, e1 = m1.eval(None, glob[0], lam)
        p3, _, e3 = sh.eval(None, per[0][rank], lam)
        n = sizes[0]
        errs = {"loss dp": rel(l2, l1), "loss sharded": rel(l3, l1),
                "eval preds sharded": rel(p3, p1[rank * n:(rank + 1) * n])}
        emb1 = m1.get_tensor("emb_mtx")
        errs["emb dp"] = rel(m2.get_tensor("emb_mtx"), emb1)
        errs["emb_m dp"] = rel(m2.get_tensor("emb_mtx/Adam"), m1.get_tensor("emb_mtx/Adam"))
        loc = parallel.global_to_local_table(torch.from_numpy(emb1), world, rank).numpy()
        errs["emb sharded"] = rel(m3.get_tensor("emb_mtx")[1:], loc[1:])
        for nm in ("fc1/kernel", "dense_3/kernel", "gru_user_side/gru_cell/gates/kernel"):
            errs["%s dp" % nm] = rel(m2.get_tensor(nm), m1.get_tensor(nm))
            errs["%s m dp" % nm] = rel(m2.get_tensor(nm + "/Adam"), m1.get_tensor(nm + "/Adam"))
            errs["%s sharded" % nm] = rel(m3.get_tensor(nm), m1.get_tensor(nm))
        # replicas must be bit-identical with each other
        t = torch.from_numpy(m2.get_tensor("emb_mtx")).cuda()
        t0 = t.clone()
        dist.broadcast(t0, 0)
        errs["replica divergence (must be 0)"] = float((t != t0).sum().item())
        if rank == 0:
            print("== adam=%s world=%d shape=%s variant=%s" % (mode, world, shape.name, variant))
            for k, v in errs.items():
                print("   %-52s %.3e" % (k, v))
        worst = max(worst, max(v for k, v in errs.items() if not k.endswith("kernel dp") and not k.endswith("kernel sharded")))
        assert errs["replica divergence (must be 0)"] == 0.0
        for m in (m1, m2, m3):
            m.close()
    ok = worst < 2e-4
    if rank == 0:
        print("MULTIGPU_CHECK", "OK" if ok else "FAILED", "worst %.3e" % worst)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
