"""torchrun script: GPU-side timeline of the data-parallel step (CUDA events on the handle's stream between the phases
of DataParallelTrainer), Taobao shape.  Usage: torchrun ... tools/dp_timeline.py [steps]"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from score_b200 import model as sb, parallel  # noqa: E402
from score_b200.synth import SHAPES, make_batch  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    shape = SHAPES["taobao"]
    m = sb.SCORE(*shape.ctor_args(), device=local, adam_mode="lazy", use_graph=True, max_batch=shape.batch)
    dp = parallel.DataParallelTrainer(m, world, rank)
    pool = [tuple(torch.from_numpy(x).cuda() for x in make_batch(shape, seed=1000 * (rank + 1) + i)) for i in range(8)]
    st = dp.stream
    names = ["begin", "count+pack", "all_gather", "finish"]
    acc = [0.0] * len(names)
    n = 0
    for i in range(steps + 10):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record(st)
        dp.begin(pool[i % 8], 5e-4, 1e-4)
        ev[1].record(st)
        cap = parallel.exchange_capacity(dp.exchange_counts(dp.local_count()))
        block = dp.pack(cap)
        ev[2].record(st)
        nw = block.numel() * world
        if dp.p2p:
            g = dp._exchange_p2p(block, cap)
        else:
            with torch.cuda.stream(st):
                if dp._gathered is None or dp._gathered.numel() < nw:
                    dp._gathered = torch.empty(nw + nw // 4, dtype=torch.int32, device=dp.device)
                g = dp._gathered[:nw]
                dist.all_gather_into_tensor(g, block)
        ev[3].record(st)
        dp.finish(g, cap, want_loss=False)
        ev[4].record(st)
        if i % 10 == 9:
            m.wait()
        if i >= 10:
            torch.cuda.synchronize()
            for k in range(4):
                acc[k] += ev[k].elapsed_time(ev[k + 1])
            n += 1
    if rank == 0:
        print("DP_TIMELINE world=%d cap=%d block_words=%d  " % (world, cap, block.numel()) +
              "  ".join("%s %.1f us" % (nm, 1e3 * a / n) for nm, a in zip(names, acc)) +
              "  total %.1f us" % (1e3 * sum(acc) / n))
    m.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
