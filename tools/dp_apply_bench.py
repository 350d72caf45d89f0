"""Single-GPU timing of the data-parallel finish (score_dp_finish) over `world` gathered blocks: one handle produces
its own block every step; the other ranks' blocks are blocks of other batches prepared up front (their contents are
what real peers would send).  Usage: python tools/dp_apply_bench.py [world] [steps]"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from score_b200 import model as sb, parallel  # noqa: E402
from score_b200.synth import SHAPES, make_batch  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    shape = SHAPES["taobao"]
    m = sb.SCORE(*shape.ctor_args(), adam_mode="lazy", use_graph=True, max_batch=shape.batch)
    dp = parallel.DataParallelTrainer(m, world, 0)
    pool = [tuple(torch.from_numpy(x).cuda() for x in make_batch(shape, seed=77 + i)) for i in range(world + 4)]
    counts = []
    for b in pool:
        dp.begin(b, 5e-4, 1e-4)
        counts.append(dp.local_count())
    cap = parallel.exchange_capacity(counts)
    foreign = []
    for b in pool[:world]:
        dp.begin(b, 5e-4, 1e-4)
        blk = dp.pack(cap)
        torch.cuda.synchronize()
        foreign.append(blk.clone())
    words = foreign[0].numel()
    gathered = torch.cat(foreign)
    st = dp.stream
    tot, n = 0.0, 0
    for i in range(steps + 5):
        dp.begin(pool[world + i % 4], 5e-4, 1e-4)
        blk = dp.pack(cap)
        with torch.cuda.stream(st):
            gathered[:words].copy_(blk)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        dp.finish(gathered, cap, want_loss=False)
        e1.record(st)
        m.wait()
        torch.cuda.synchronize()
        if i >= 5:
            tot += e0.elapsed_time(e1); n += 1
    print("DP_FINISH world=%d cap=%d unique/rank~%d: %.1f us per finish" % (world, cap, counts[0], 1e3 * tot / n))
    m.close()


if __name__ == "__main__":
    main()
