"""GPU timeline of the row-sharded step (torchrun, N ranks): CUDA events at the phase boundaries of
ShardedEmbeddingTrainer.train_async, averaged over the timed steps, rank 0 prints one SHARD_TIMELINE line.
Usage: torchrun ... tools/shard_timeline.py [workload] [steps]"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from score_b200 import model as sb
from score_b200 import parallel
from score_b200.synth import SHAPES, make_batch


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shape = SHAPES[sys.argv[1] if len(sys.argv) > 1 else "large_vocab"]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    ctor = list(shape.ctor_args())
    ctor[0] = parallel.shard_rows(shape.feature_size, world)
    m = sb.SCORE(*ctor, device=local, adam_mode="lazy", use_graph=True, seed=1111, max_batch=shape.batch)
    tr = parallel.ShardedEmbeddingTrainer(m, world, rank)
    pool = [tuple(torch.from_numpy(x).cuda() for x in make_batch(shape, seed=7000 * (rank + 1) + i)) for i in range(8)]
    for i in range(6):
        tr.train_async(pool[i % 8], 5e-4, 1e-4)
    tr.wait()
    dist.barrier(); torch.cuda.synchronize()
    tr.timeline = {}
    for i in range(steps):
        tr.train_async(pool[i % 8], 5e-4, 1e-4)
    tr.wait()
    torch.cuda.synchronize()
    tl = tr.timeline
    names = ["start", "plan+counts", "finish(prev)", "a2a ids", "gather", "a2a rows", "fwd+bwd", "allreduce", "pack", "a2a grads"]
    out = []
    for a, b in zip(names[:-1], names[1:]):
        ms = [x.elapsed_time(y) for x, y in zip(tl[a], tl[b])]
        out.append("%s %.0f" % (b, 1e3 * sum(ms) / len(ms)))
    gap = [tl["a2a grads"][i].elapsed_time(tl["start"][i + 1]) for i in range(steps - 1)]
    total = tl["start"][0].elapsed_time(tl["start"][-1]) / (steps - 1)
    if rank == 0:
        print("SHARD_TIMELINE world=%d %s: step %.0f us | %s | host gap to next start %.0f  (us; 'a2a ids' includes the host's wait for the counts)"
              % (world, shape.name, 1e3 * total, " / ".join(out), 1e3 * sum(gap) / len(gap)))
    m.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
