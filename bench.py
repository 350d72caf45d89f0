#!/usr/bin/env python
"""bench.py - SCoRe training throughput (samples/s) on synthetic data of the reference's shapes.

One "step" = one pass of the hot path (forward + backward + Adam, code/score/score.py:101-116) over one
synthetic batch.  Workload at N=1: Taobao shape (BASELINE.json configs[2]: V=5 042 754 rows, d=16, T=8,
K=10, batch 1024), the config the >=100x target is quoted on.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload taobao]

Prints ONE JSON line (rank 0).  `value` = whole-job samples/s with the id tensors already resident in
HBM; `e2e` = the same metric through the reference-facing call model.train(sess, batch, lr, reg) with
pinned HOST buffers (H2D of the ids and D2H of the loss inside the timed region).  `roofline` is the
dominant HBM-bound kernel of the step - the scatter's segment-reduce + fused row Adam (emb_update) - timed with
CUDA events on its own stream inside the timed steps; `roofline_gather` is the fused embedding gather +
co-attention kernel (coatt_fwd) measured the same way.  `cpu_baseline` is the
literal CPU restatement of score.py (oracle/, "port": TF 1.x cannot be installed offline) timed on the
box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LR, REG = 5e-4, 1e-4          # train_score.py:371-372 (first grid point)
POOL = 16                      # distinct batches rotated through the timed region


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json, written by
    tools/ncu_traffic.py); None when no capture of this workload exists."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        e = json.load(open(path))[workload][kernel]
        return e["traffic_bytes"], e["source"]
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class LineGuard(object):
    """stdout carries exactly ONE JSON line, whatever happens to the optional tail of the run: emit() writes once; between
    arm(line) and disarm() the rescue timer may close the run - rank 0 prints the measured line with the unfinished leg
    marked, every rank leaves with status 0 (a stuck collective in the extra leg must not cost the main measurement)"""

    def __init__(self, fd, rank, limit_s, exit_fn=os._exit):
        self.fd, self.rank, self.limit_s, self.exit_fn = fd, rank, limit_s, exit_fn
        self.lock = threading.Lock()
        self.done = False
        self.armed = None

    def emit(self, line):
        with self.lock:
            if not self.done:
                self.done = True
                os.write(self.fd, (json.dumps(line) + "\n").encode())

    def arm(self, line):
        self.armed = line if self.rank == 0 else {}

    def disarm(self):
        self.armed = None

    def rescue(self):
        line = self.armed
        if line is None or self.done:
            return
        if self.rank == 0:
            self.emit(dict(line, large_vocab={"error": "not finished within the bench's time limit (%d s)" % self.limit_s}))
        with self.lock:
            self.done = True
        self.exit_fn(0)

    def start_timer(self, seconds):
        t = threading.Timer(seconds, self.rescue)
        t.daemon = True
        t.start()
        return t


def cpu_baseline(shape, batch_size, budget_s=20.0, max_steps=8):
    """Literal CPU restatement of score.py timed on the host cores (bounded sample of the workload)."""
    import torch
    from oracle import score_ref as ref
    from score_b200.synth import make_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    o = ref.ScoreOracle(*shape.ctor_args(), seed=1111)
    b = make_batch(shape, batch=batch_size, seed=1)
    o.train(None, b, LR, REG)                      # warm-up (allocations, thread pools)
    times = []
    t_all = time.perf_counter()
    for i in range(max_steps):
        b = make_batch(shape, batch=batch_size, seed=2 + i)
        t0 = time.perf_counter()
        o.train(None, b, LR, REG)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget_s:
            break
    med = float(np.median(times))
    return {"value": batch_size / med, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "%d train steps of batch %d (median step %.3f s); literal restatement of score.py in "
                      "PyTorch-CPU fp32 (materialised KxK co-attention, dense-table Adam); TF 1.x could not be "
                      "installed offline" % (len(times), batch_size, med)}, med


def run_reference(args, shape, rank, world, emit):
    if rank != 0:
        return
    cb, med = cpu_baseline(shape, shape.batch, budget_s=max(10.0, 3.0 * (args.steps + args.warmup)),
                           max_steps=max(1, args.steps))
    line = {"impl": "reference", "metric": "train_samples_per_sec", "value": cb["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the arm's config at this N (global batch = per-GPU batch x N); a step of THIS arm is a bounded sample of it:
            # one batch of `per_gpu_batch` samples on the host cores (cpu_baseline.sample) - samples/s does not depend on it
            "config": workload_config(shape, args, shape.batch * max(1, args.gpus)),
            "cpu_baseline": dict(cb),
            "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_config(shape, args, global_batch):
    return {"workload": "SCoRe %s-shape synthetic (V=%d rows, d=%d, T=%d, K=%d, uf/if=%d/%d, H=%d)" % (
                shape.name, shape.feature_size, shape.eb_dim, shape.max_time_len, shape.obj_per_time_slice,
                shape.user_fnum, shape.item_fnum, shape.hidden_size),
            "global_batch": global_batch, "per_gpu_batch": shape.batch, "length": shape.length,
            "ids": "uniform" if not args.zipf else "zipf %.2f" % args.zipf,
            "adam": args.adam_mode, "cuda_graph": not args.no_graph,
            # the same text in both arms of a run at this N (the reference arm cannot know which scheme the probe picks):
            # the scheme that ran is the line's top-level `parallelism_scheme`
            "parallelism": "single GPU" if args.gpus <= 1 else
                           "dp%d: batch split over %d GPUs (%d samples each); embedding table replicated (one packed all-gather per "
                           "step) or row-sharded by id %% %d (peer-memory exchange over NVLink), as selected for this N - see "
                           "parallelism_scheme" % (args.gpus, args.gpus, shape.batch, args.gpus),
            "l2_policy": "inputs larger than L2: %.2f GB of embedding state (var+m+v), uniform-random rows, %d rotating "
                         "batches; no explicit flush" % (3 * shape.feature_size * shape.eb_dim * 4 / 1e9, POOL)}


def scheme_info(args, par_probe):
    """which multi-GPU scheme the timed region ran (and the probe that chose it)"""
    text = {"single": "single GPU",
            "dp": "dp%d: replicated table, ONE all-gather per step of packed blocks (dense gradient + one embedding-"
                  "gradient row per unique id), rank-ordered deterministic reduce on every replica" % args.gpus,
            "sharded": "dp%d dense + embedding rows sharded by id %% %d: ids through one NCCL all-to-all, rows and "
                       "gradient rows stored straight into the peers' buffers over NVLink (peer memory; "
                       "SCORE_SHARD_P2P=0: NCCL all-to-alls)" % (args.gpus, args.gpus)}[getattr(args, "par", "single")]
    return {"chosen": getattr(args, "par", "single"), "description": text, "probe_ms_per_step": par_probe}


def rooflines(shape, stats, probes):
    """(roofline_gather, roofline_scatter) of the bench line from the step statistics and the per-kernel event times
    (kept apart from main() so the CPU tests can run it)"""
    peak, peak_src = peaks()
    d = shape.eb_dim
    live, uniq = stats["live"], stats["unique_rows"]
    gather_bytes = live * (4 + 4 * d)
    scatter_bytes = live * (4 + 4 * d) + uniq * 6 * 4 * d

    def per_launch(name):
        tot, n = probes.get(name, (0.0, 0))
        return tot / n if n else None

    t_g, t_s = per_launch("coatt_fwd"), per_launch("emb_update")
    tr_g, src_g = ncu_traffic(shape.name, "coatt_fwd")
    tr_s, src_s = ncu_traffic(shape.name, "emb_update")
    roof = {"bound": "hbm", "kernel": "coatt_fwd_lean_kernel / coatt_fwd_kernel (fused embedding gather + co-attention + pooling; the lean instance serves the compiled-in geometries, Taobao among them)",
            "achieved": gather_bytes / (t_g * 1e-3) / 1e9 if t_g else None, "peak": peak, "unit": "GB/s",
            "frac": gather_bytes / (t_g * 1e-3) / 1e9 / peak if t_g else None, "traffic": tr_g,
            "traffic_source": src_g,
            "peak_source": peak_src, "bytes_per_launch": gather_bytes, "ms_per_launch": t_g,
            "bytes_rule": "live non-zero ids x (4 + 4d); every index counted, no credit for the duplicates the "
                          "loader's cyclic padding and the 1+neg user-side replication create (those hit L2, "
                          "which is why traffic < bytes_per_launch)",
            "timing": "CUDA events around the kernel on its own stream inside every step of a second timed region of "
                      "the same K steps (ms_per_step_probed); the sort / weight-gradient streams run concurrently"}
    roof["random_access_ceiling"] = {
        "frac_at_this_launch_size": 0.28, "frac_asymptotic": 0.45,
        "what": "a kernel that ONLY gathers the same number of random 64-byte rows (129.5 B of DRAM reads per row on this part)",
        "source": "tools/randrow_bench.cu, profiles/r2d_randrow_bench.txt (d = 16 tables)"} if d == 16 else None
    roof_s = {"bound": "hbm", "kernel": "emb_update_kernel (segment-reduce + fused row Adam)",
              "achieved": scatter_bytes / (t_s * 1e-3) / 1e9 if t_s else None, "peak": peak, "unit": "GB/s",
              "frac": scatter_bytes / (t_s * 1e-3) / 1e9 / peak if t_s else None, "traffic": tr_s,
              "traffic_source": src_s, "peak_source": peak_src,
              "timing": "CUDA events around the kernel on its own stream inside every step of a second timed region "
                        "of the same K steps (ms_per_step_probed; the event nodes add ~3 us to each bracketed kernel)",
              "bytes_per_launch": scatter_bytes, "ms_per_launch": t_s, "unique_rows": uniq,
              "random_access_ceiling": {"frac": 0.41, "what": "a kernel that ONLY reads and rewrites the same number of "
                                                               "random 192-byte records (var | m | v)",
                                        "source": "tools/randrow_bench.cu, profiles/r2d_randrow_bench.txt (d = 16 tables)"}
              if d == 16 else None,
              "bytes_rule": "live ids x (4 + 4d) + unique rows x 6 x 4d"}
    return roof, roof_s


def large_vocab_leg(args, world, rank, local, steps=30, warmup=5):
    """N > 1 only: BASELINE.json config 5 (200 M-row table, d=64, row-sharded over the N GPUs, batch 1024 per GPU) and,
    beside it, ONE GPU stepping on a 25 M-row shard of the same table (the per-GPU share at 8 GPUs) - the denominator of
    the >= 6x target.  Same timing rules as the main line: barrier + synchronize on both sides, CUDA events on the
    launching stream, max over ranks."""
    import torch
    import torch.distributed as dist
    from score_b200 import model as sb
    from score_b200 import parallel
    from score_b200.synth import SHAPES, make_batch
    out = {}
    dev = torch.device("cuda", local)

    def run(shape, sharded):
        ctor = list(shape.ctor_args())
        if sharded:
            ctor[0] = parallel.shard_rows(shape.feature_size, world)
        m = sb.SCORE(*ctor, device=local, adam_mode=args.adam_mode, use_graph=not args.no_graph, seed=1111, max_batch=shape.batch)
        trainer = parallel.ShardedEmbeddingTrainer(m, world, rank) if sharded else None
        stream = torch.cuda.ExternalStream(m.stream(), device=dev)
        # one distinct batch per step: with uniform ids over 10^8 rows a row practically never comes back, a short
        # rotating pool would make EVERY row of a batch stale by pool-size steps and bill that replay to the step
        npool = steps + warmup + 1
        pool = [tuple(torch.from_numpy(x).cuda() for x in make_batch(shape, seed=7000 * (rank + 1) + i)) for i in range(npool)]

        def step(i):
            (trainer or m).train_async(pool[i % npool], LR, REG)

        for i in range(warmup):
            step(i)
        (trainer or m).wait()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for i in range(steps):
                step(warmup + i)
            if trainer:
                trainer._flush_finish()       # the last step's deferred optimizer half belongs to the timed region
            e1.record(stream)
        (trainer or m).wait()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        loss = (trainer.train(None, pool[-1], LR, REG) if trainer else m.train(None, pool[-1], LR, REG))
        m.close()
        del pool, m, trainer
        torch.cuda.empty_cache()
        return float(t.item()) / steps, loss

    lv = SHAPES["large_vocab"]
    ms, loss = run(lv, True)
    out.update({"workload": "SCoRe large_vocab synthetic (V=%d rows, d=%d, H=%d), rows sharded by id %% %d over %d GPUs" % (
                    lv.feature_size, lv.eb_dim, lv.hidden_size, world, world),
                "value": lv.batch * world / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "per_gpu_batch": lv.batch,
                "steps": steps, "warmup": warmup, "loss": loss})
    one = SHAPES["large_vocab_shard"]
    ms1, _ = run(one, False)     # every rank runs its own replica (no collective); the slowest one counts
    v1 = one.batch / (ms1 * 1e-3)
    out["one_gpu_shard"] = {"workload": "the same step on ONE GPU holding a %d-row shard (the per-GPU share at 8 GPUs), batch %d"
                            % (one.feature_size, one.batch), "value": v1, "ms_per_step": ms1}
    out["x_vs_1gpu_shard"] = out["value"] / v1
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="taobao")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's own, BASELINE.json)")
    ap.add_argument("--adam-mode", default="lazy", choices=["dense", "lazy", "sparse"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--zipf", type=float, default=0.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-probes", action="store_true", help="experiment: timed region without the per-kernel event probes")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--parallel", default="auto", choices=["auto", "dp", "sharded"],
                    help="N>1: dp = replicated table + gradient all-gather, sharded = row-sharded table + all-to-all")
    ap.add_argument("--no-large-vocab", action="store_true",
                    help="N>1: skip the extra large_vocab leg (row-sharded 200 M-row table + the 1-GPU 25 M-row shard)")
    ap.add_argument("--watchdog", type=int, default=env_int("SCORE_BENCH_WATCHDOG", 480),
                    help="seconds after which every thread's stack is dumped to stderr and the process exits (a stuck "
                         "collective must not hang the caller); 0 disables")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.watchdog > 0 and args.impl == "ours":
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog, exit=True)
    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner) go to stderr instead
    json_fd = os.dup(1)
    os.dup2(2, 1)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    guard = LineGuard(json_fd, rank, args.watchdog)
    emit = guard.emit
    if args.watchdog > 60 and args.impl == "ours":
        guard.start_timer(args.watchdog - 30)     # 30 s before the hard watchdog

    from score_b200.synth import SHAPES, make_batch
    shape = SHAPES[args.workload]
    if args.batch > 0:
        import dataclasses
        shape = dataclasses.replace(shape, batch=args.batch)
    if args.impl == "reference":
        run_reference(args, shape, rank, world, emit)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from score_b200 import model as sb
    from score_b200 import parallel

    B = shape.batch
    host_pool = [make_batch(shape, batch=B, seed=1000 * (rank + 1) + i, zipf=args.zipf) for i in range(POOL)]
    dev_pool = [tuple(torch.from_numpy(x).cuda() for x in b) for b in host_pool]

    def build(par_):
        ctor = list(shape.ctor_args())
        if world > 1 and par_ == "sharded":
            ctor[0] = parallel.shard_rows(shape.feature_size, world)
        # every rank draws the same dense initial values (the dense part is data-parallel: replicas must start equal); a
        # row-sharded table is drawn per rank from the same stream, which only makes the shards look alike - harmless here
        m_ = sb.SCORE(*ctor, device=local, adam_mode=args.adam_mode, use_graph=not args.no_graph, seed=1111, max_batch=B)
        t_ = None
        if world > 1:
            t_ = (parallel.ShardedEmbeddingTrainer if par_ == "sharded" else parallel.DataParallelTrainer)(m_, world, rank)
        return m_, t_

    # Multi-GPU on a table that fits one GPU: both schemes apply - replicated table + packed all-gather, or row-sharded
    # table + all-to-all - and which one wins depends on N (the replicated update grows with the world size, the sharded
    # exchange has a fixed latency): a short probe of both picks the faster one for this N (recorded in config)
    par_probe = None
    par = args.parallel
    if par == "auto":
        par = "sharded" if shape.feature_size > 50_000_000 else "dp"
        if world > 1 and par == "dp":
            par_probe = {}
            for cand in ("dp", "sharded"):
                m_, t_ = build(cand)
                for i in range(8):
                    t_.train_async(dev_pool[i % POOL], LR, REG)
                t_.wait()
                dist.barrier(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(30):
                    t_.train_async(dev_pool[i % POOL], LR, REG)
                t_.wait()
                torch.cuda.synchronize()
                tt = torch.tensor([(time.perf_counter() - t0) / 30 * 1e3], device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                par_probe[cand] = float(tt.item())
                m_.close()
                del m_, t_
                torch.cuda.empty_cache()
            par = min(par_probe, key=par_probe.get)
    args.par = par if world > 1 else "single"
    m, trainer = build(par)
    stream = torch.cuda.ExternalStream(m.stream(), device=torch.device("cuda", local))

    pin_pool = [tuple(torch.from_numpy(x).pin_memory() for x in b) for b in host_pool]
    torch.cuda.synchronize()

    def step_dev(i):
        b = dev_pool[i % POOL]
        if trainer:
            trainer.train_async(b, LR, REG)
        else:
            m.train_async(b, LR, REG)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput
    # clocks / throttle reasons are sampled from here to the end of the timed regions (nvidia-smi needs ~100 ms to
    # deliver its first line, the default timed region is shorter than that): warm-up and timed steps are the same load
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    def wait_all():
        return trainer.wait() if trainer else m.wait()

    for i in range(args.warmup):
        step_dev(i)
    wait_all()
    m.enable_probes(False)

    def timed_region(steps):
        """EXACTLY `steps` steps between barrier + synchronize, CUDA events on the launching stream"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for i in range(steps):
                step_dev(i)
            if trainer and hasattr(trainer, "_flush_finish"):
                trainer._flush_finish()       # row-sharded: the last step's deferred optimizer half is part of the region
            e1.record(stream)
        last = wait_all()
        barrier()
        t_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([t_ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
        return t_ms, last

    # (1) the throughput region: no instrumentation inside the step
    launches0 = m.launch_count()
    ms, loss = timed_region(args.steps)
    launches = m.launch_count() - launches0
    ms_per_step = ms / args.steps
    value = B * world / (ms_per_step * 1e-3)
    # (2) the same region again with the per-kernel event probes recorded inside every step (the graphs are re-captured
    #     with the event nodes: 16 extra nodes per step, measured at +7 % on the step) -> kernel durations for the rooflines
    probes, ms_probed = {}, None
    if not args.no_probes:
        m.enable_probes(True)
        for i in range(3):
            step_dev(i)
            wait_all()
        m.enable_probes(True)       # same state: resets the accumulators only, graphs are kept
        ms_p, _ = timed_region(args.steps)
        ms_probed = ms_p / args.steps
        probes = m.probe_times()
    # very short runs: keep the load up until a few clock samples exist.  The steps are collective at N > 1, so rank 0
    # (which owns the sampler) decides and every rank follows the broadcast decision
    for _ in range(10):
        need = 1 if (rank == 0 and len(sampler.rows) < 3) else 0
        if world > 1:
            flag = torch.tensor([need], device="cuda")
            dist.broadcast(flag, 0)
            need = int(flag.item())
        if not need:
            break
        for i in range(40):
            step_dev(i)
        wait_all()
    clocks = sampler.stop() if rank == 0 else None
    stats = m.last_step_stats()

    # ---------------- end to end through the reference-facing API, host buffers
    m.enable_probes(False)
    e2e_steps = args.e2e_steps or max(20, args.steps // 4)

    def step_host(i):
        b = pin_pool[i % POOL]
        return trainer.train(None, b, LR, REG) if trainer else m.train(None, b, LR, REG)

    for i in range(3):
        step_host(i)
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        ev2.record(stream)
        for i in range(e2e_steps):
            loss_e2e = step_host(i)
        ev3.record(stream)
    torch.cuda.synchronize()
    e2e_ms = max(ev2.elapsed_time(ev3), (time.perf_counter() - t0) * 1e3)
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = B * world / (e2e_ms / e2e_steps * 1e-3)
    h2d = int(sum(x.numel() * 4 for x in pin_pool[0]))
    # the reference's caller feeds NESTED PYTHON LISTS (graph_loader.py:383 -> score.py:102-115, TF converts them inside
    # sess.run): the same call with lists, so the list -> int32 array conversion the boundary must do is inside the region
    e2e_lists = None
    if world == 1:
        list_pool = [tuple(x.tolist() for x in b) for b in host_pool[:4]]
        m.train(None, list_pool[0], LR, REG)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_l = 6
        for i in range(n_l):
            m.train(None, list_pool[i % 4], LR, REG)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_l
        e2e_lists = {"value": B / dt, "unit": "samples/s", "ms_per_step": dt * 1e3, "steps": n_l,
                     "note": "batch_data as nested Python lists, as GraphLoader yields them: the step is the same, the time is "
                             "the walk over %d Python ints (+ their lists) per batch on one host core (csrc/listfeed.c)" % (h2d // 4)}
        del list_pool
    if trainer:
        loss = loss_e2e      # the trainer's synchronous step returns the GLOBAL loss (m.wait() is this rank's share only)
    m.close()
    trainer = None       # releases the trainer's exchange buffers (symmetric memory) while the process group is alive
    del dev_pool
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    line = None
    if rank == 0:
        roof, roof_s = rooflines(shape, stats, probes)
        line = {"metric": "train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(shape, args, B * world),
                "parallelism_scheme": scheme_info(args, par_probe),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 16 if world == 1 else 20,
                        "steps": e2e_steps,
                        "api": "SCORE.train(sess, batch_data, lr, reg_lambda) -> float, pinned host ids; the call returns "
                               "when the step's result packet (loss, error flag) has reached the host, the region ends "
                               "with a device synchronize",
                        "nested_lists": e2e_lists},
                "gpu_launches": int(launches),
                # `roofline` = the dominant HBM-bound kernel of the step (most DRAM traffic, longest of the HBM-side
                # kernels): the scatter + row Adam; the gather kernel is reported beside it (BASELINE.json's metric
                # names both)
                "roofline": roof_s, "roofline_gather": roof, "roofline_scatter": roof_s,
                "kernel_ms": {k: (v[0] / v[1] if v[1] else None) for k, v in probes.items()},
                "ms_per_step_probed": ms_probed,
                "final_loss": loss}
    if world > 1 and not args.no_large_vocab and args.workload == "taobao":
        guard.arm(line)      # from here on the rescue timer may close the run
        try:
            lv_leg = large_vocab_leg(args, world, rank, local)
        except Exception as e:      # the main line must survive a failure of the extra leg
            lv_leg = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        guard.disarm()
        if rank == 0:
            line["large_vocab"] = lv_leg
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            cb, _ = cpu_baseline(shape, B)
            line["cpu_baseline"] = cb
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
