"""Row-sharded step with ONE rank (no process group: the exchanges degenerate to copies) for single-GPU profiling of
the owner-side kernels.  Usage: python tools/shard_world1.py [workload] [steps]"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from score_b200 import model as sb
from score_b200 import parallel
from score_b200.synth import SHAPES, make_batch

shape = SHAPES[sys.argv[1] if len(sys.argv) > 1 else "large_vocab_shard"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctor = list(shape.ctor_args())
ctor[0] = parallel.shard_rows(shape.feature_size, 1)
m = sb.SCORE(*ctor, adam_mode="lazy", use_graph=False, seed=1111, max_batch=shape.batch)
tr = parallel.ShardedEmbeddingTrainer(m, 1, 0)
pool = [tuple(torch.from_numpy(x).cuda() for x in make_batch(shape, seed=7000 + i)) for i in range(4)]
for i in range(steps):
    tr.train_async(pool[i % 4], 5e-4, 1e-4)
tr.wait()
torch.cuda.synchronize()
print("SHARD_WORLD1_OK")
m.close()
