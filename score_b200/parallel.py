"""Multi-GPU stepping of the SCoRe hot path: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch) for the exchanges, the C-ABI split step (score_step_begin / score_step_finish) for the math.

The reference is single-device; both schemes keep SCOREBASE.train's arithmetic
(score.py:101-116) on the GLOBAL batch:

* ``DataParallelTrainer`` - replicated table (Tmall / Taobao / CCMR sizes).  Samples are independent (BN runs in
  inference mode, the loss is a batch mean), so each rank runs forward/backward on its shard with 1/global_B
  scaling; dense gradients are summed with ONE all-reduce of one flat buffer; each rank reduces its embedding
  gradient to one row per unique id (deterministic segment reduce), those (key, row) lists are all-gathered and
  every replica applies the SAME deterministic sort + segment-reduce (rank order) + Adam, so replicas stay
  bit-identical without ever broadcasting the table.

* ``ShardedEmbeddingTrainer`` - row-sharded table (large-vocab config): owner(id) = id % world, local row
  = id // world + 1 (local row 0 is the dummy).  Forward: positions grouped by owner on the device (score_shard_plan)
  -> all-to-all ids -> the owners bring the rows up to date (lazy Adam) and store them straight into the requesters'
  staged tables over NVLink (peer memory) -> the unchanged kernels run on the staged table.  Backward: the gradient rows
  are stored straight into the owners' buffers -> owner-side sort (done earlier, on a side stream) / reduce / Adam.

The pure-torch exchange helpers below run on CPU tensors with the gloo backend too (tests/test_parallel_cpu.py).
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

from . import _capi
from .model import TRAIN_KEEP_PROB, _Batch


# ------------------------------------------------------------------------------------------ exchange helpers
class ExchangePlan:
    """How one rank's flat position list is bucketed by owner for the all-to-all."""

    def __init__(self, keys: torch.Tensor, world: int, group=None):
        n = keys.numel()
        k64 = keys.to(torch.int64)
        nz = k64 != 0
        owner = torch.where(nz, k64 % world, torch.full_like(k64, world))   # dummy positions -> bucket `world`
        self.order = torch.argsort(owner, stable=True)                       # grouped by owner, position order inside
        counts = torch.bincount(owner, minlength=world + 1)[:world]
        recv = torch.empty_like(counts)
        dist.all_to_all_single(recv, counts, group=group)
        self.send_counts = counts.tolist()                                   # host sync: sizes of the exchange
        self.recv_counts = recv.tolist()
        self.n_valid = int(sum(self.send_counts))
        self.n_recv = int(sum(self.recv_counts))
        self.sel = self.order[:self.n_valid]                                 # positions that have a real row
        self.send_local_rows = (k64[self.sel] // world + 1).to(torch.int32)  # owner-local row numbers
        self.mini_keys = torch.where(nz, torch.arange(1, n + 1, device=keys.device), torch.zeros_like(k64)).to(torch.int32)
        self.world, self.group, self.n = world, group, n

    def exchange_ids(self):
        """-> local row numbers this rank must serve, grouped by requesting rank."""
        out = torch.empty(self.n_recv, dtype=torch.int32, device=self.send_local_rows.device)
        dist.all_to_all_single(out, self.send_local_rows, self.recv_counts, self.send_counts, group=self.group)
        return out

    def return_rows(self, served_rows: torch.Tensor):
        """owner -> requester: rows for `exchange_ids()` order come back in `sel` order; builds the staged table."""
        d = served_rows.shape[1]
        got = torch.empty(self.n_valid, d, dtype=served_rows.dtype, device=served_rows.device)
        dist.all_to_all_single(got, served_rows, self.send_counts, self.recv_counts, group=self.group)
        staged = torch.zeros(self.n + 1, d, dtype=served_rows.dtype, device=served_rows.device)
        staged[self.sel + 1] = got
        return staged

    def send_grads(self, grad_rows: torch.Tensor):
        """requester -> owner: per-position gradient rows, aligned with `exchange_ids()`."""
        d = grad_rows.shape[1]
        out = torch.empty(self.n_recv, d, dtype=grad_rows.dtype, device=grad_rows.device)
        dist.all_to_all_single(out, grad_rows[self.sel].contiguous(), self.recv_counts, self.send_counts, group=self.group)
        return out


def shard_rows(n_rows_global: int, world: int) -> int:
    """rows of one rank's local table (incl. the local dummy row 0)."""
    return (n_rows_global + world - 1) // world + 1


def global_to_local_table(table: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    """slice a global [V,d] table into rank's local layout (row 1 + v // world for v % world == rank)."""
    mine = table[rank::world]
    out = torch.zeros(shard_rows(table.shape[0], world), table.shape[1], dtype=table.dtype)
    out[1:1 + mine.shape[0]] = mine
    return out


# ------------------------------------------------------------------------------------------ device plumbing
class _DevView:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class _Base:
    def __init__(self, model, world, rank, group=None):
        self.m, self.world, self.rank, self.group = model, int(world), int(rank), group
        self.lib = model._lib
        self.h = model._h
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.ExternalStream(model.stream(), device=self.device)

    def _dev(self, name, dtype):
        ptr, cnt = C.c_void_p(), C.c_size_t()
        self.m._check(self.lib.score_device_buffer(self.h, name.encode(), C.byref(ptr), C.byref(cnt)))
        if cnt.value == 0:
            return torch.empty(0, dtype=dtype, device=self.device)
        ts = "<f4" if dtype == torch.float32 else "<i4"
        return torch.as_tensor(_DevView(ptr.value, (cnt.value,), ts), device=self.device)

    def _global_loss(self, loss2):
        """sum over ranks of the data term (already scaled by 1/global_B) + the L2 term once."""
        t = torch.tensor([loss2[0] - loss2[1]], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, group=self.group)
        return float(t.item()) + float(loss2[1])


def split_count_matrix(cm, world: int, rank: int):
    """cm: the all-gathered count matrix, flat [world][world + 1]: cm[r][o] = positions of rank r owned by rank o, column
    `world` = rank r's dummy positions (id 0).  -> (send_counts, recv_counts) of `rank` for the all-to-alls, and two numbers
    that are identical on every rank (all of them hold the whole matrix) and size the symmetric exchange buffers: the
    largest number of rows any owner serves, the largest position count of any rank."""
    W = world
    send_counts = cm[rank * (W + 1):rank * (W + 1) + W]
    recv_counts = [cm[r * (W + 1) + rank] for r in range(W)]
    max_recv = max(sum(cm[r * (W + 1) + o] for r in range(W)) for o in range(W))
    max_pos = max(sum(cm[r * (W + 1):(r + 1) * (W + 1)]) for r in range(W))
    return send_counts, recv_counts, max_recv, max_pos


def exchange_capacity(counts) -> int:
    """list capacity of the packed exchange: the largest rank count rounded up to a multiple of 1024 (same on all ranks)."""
    return max(1024, (int(max(counts)) + 1023) // 1024 * 1024)


class DataParallelTrainer(_Base):
    """Dense data-parallel training with a replicated embedding table (BASELINE.json config 4).

    One step = score_step_begin (forward/backward on the local shard, CUDA graph) -> ONE all-gather of one packed block
    per rank (dense gradient + one embedding-gradient row per unique id, ascending) -> score_dp_finish (rank-ordered
    sums, merge of the sorted lists, row Adam).  The list capacity comes from the unique-row counts, which the sort
    branch publishes long before forward/backward ends: they are exchanged on a side stream while the main stream keeps
    running, so the host never drains the device inside a step.  The phases are separate methods so a single-process
    test can drive several handles through them (tests/test_parity_gpu.py)."""

    def __init__(self, model, world, rank, group=None, p2p=None):
        super().__init__(model, world, rank, group)
        self.side = torch.cuda.Stream(device=self.device)
        self._gathered = None
        # Opt-in (SCORE_DP_P2P=1): peer-memory exchange instead of the NCCL all-gather - every rank stores its block
        # straight into all replicas' gathered buffers (torch symmetric memory supplies the peer mappings and the
        # device-side barrier).  Measured on B200 (Taobao shape, parity as the all-gather path at 2 and 8 GPUs): 0.510 ->
        # 0.489 ms at 2 GPUs, 0.927 -> 0.864 ms at 8; the all-gather stays the default (the replicated update, not the
        # exchange, is what bounds this scheme - DESIGN.md section 5).
        self.p2p = (os.environ.get("SCORE_DP_P2P") == "1") if p2p is None else bool(p2p)
        self._symm = None
        self._step_no = 0

    def _p2p_setup(self, words):
        """two gathered buffers (alternating steps) of world blocks each in symmetric memory, mapped into every rank."""
        import torch.distributed._symmetric_memory as symm_mem
        n_pos = self._dev("keys", torch.int32).numel()
        cap_max = exchange_capacity([n_pos])                       # every position a distinct row: the hard upper bound
        self._max_words = int(self.lib.score_dp_block_words(self.h, cap_max))
        self._symm = symm_mem.empty(2 * self.world * self._max_words, dtype=torch.int32, device=self.device)
        grp = self.group if self.group is not None else dist.group.WORLD
        self._symm_hdl = symm_mem.rendezvous(self._symm, grp)
        self._peer_bases = (C.c_uint64 * self.world)(*[int(p) for p in self._symm_hdl.buffer_ptrs])

    def _exchange_p2p(self, block, cap):
        words = block.numel()
        if self._symm is None:
            self._p2p_setup(words)
        if words > self._max_words:
            raise RuntimeError("packed block larger than the symmetric buffer")
        base = (self._step_no & 1) * self.world * self._max_words     # alternate buffers: a peer may still read the other
        self._step_no += 1
        with torch.cuda.stream(self.stream):
            self.m._check(self.lib.score_dp_push(self.h, cap, self._peer_bases, self.world, base + self.rank * words))
            self._symm_hdl.barrier(channel=0)                         # all blocks stored and visible; orders buffer reuse
        return self._symm[base:base + self.world * words]

    # ---- phases
    def begin(self, batch_data, lr, reg_lambda, keep_prob=TRAIN_KEEP_PROB, global_batch=None):
        b = _Batch(batch_data, self.m.cfg)
        gb = global_batch if global_batch else b.B * self.world
        self.m._check(self.lib.score_set_sample_offset(self.h, self.rank * b.B))   # dropout masks of the global batch
        with torch.cuda.stream(self.stream):
            self.m._check(self.lib.score_step_begin(self.h, C.byref(b.struct), lr, reg_lambda, keep_prob, gb, 1, None, None))
        self._keep_batch = b   # host id arrays stay alive until their H2D copies have run
        return b

    def local_count(self) -> int:
        c = C.c_int32()
        self.m._check(self.lib.score_dp_local_count(self.h, C.byref(c)))
        return int(c.value)

    def exchange_counts(self, count):
        """all ranks' unique-row counts (NCCL on a side stream: the main stream is neither waited for nor blocked)."""
        with torch.cuda.stream(self.side):
            mine = torch.tensor([count], dtype=torch.int32, device=self.device)
            out = torch.empty(self.world, dtype=torch.int32, device=self.device)
            dist.all_gather_into_tensor(out, mine, group=self.group)
            return out.tolist()

    def pack(self, cap):
        """-> this rank's packed block as an int32 tensor view (device memory owned by the handle)."""
        ptr, words = C.c_void_p(), C.c_int64()
        with torch.cuda.stream(self.stream):
            self.m._check(self.lib.score_dp_pack(self.h, cap, C.byref(ptr), C.byref(words)))
        return torch.as_tensor(_DevView(ptr.value, (words.value,), "<i4"), device=self.device)

    def finish(self, gathered, cap, want_loss=True):
        loss = C.c_double() if want_loss else None
        with torch.cuda.stream(self.stream):
            self.m._check(self.lib.score_dp_finish(self.h, gathered.data_ptr(), self.world, cap,
                                                   C.byref(loss) if want_loss else None))
        return float(loss.value) if want_loss else None

    # ---- the step
    def train(self, sess, batch_data, lr, reg_lambda, keep_prob=TRAIN_KEEP_PROB, want_loss=True):
        self.begin(batch_data, lr, reg_lambda, keep_prob)
        return self._rest(want_loss)

    def _rest(self, want_loss):
        cap = exchange_capacity(self.exchange_counts(self.local_count()))
        block = self.pack(cap)
        if self.p2p:
            return self.finish(self._exchange_p2p(block, cap), cap, want_loss)
        n = block.numel() * self.world
        with torch.cuda.stream(self.stream):
            if self._gathered is None or self._gathered.numel() < n:
                self.stream.synchronize()          # the previous step's update may still be reading the old buffer
                self._gathered = torch.empty(n + n // 4, dtype=torch.int32, device=self.device)
            gathered = self._gathered[:n]
            dist.all_gather_into_tensor(gathered, block, group=self.group)
        return self.finish(gathered, cap, want_loss)

    def train_async(self, batch_data, lr, reg_lambda, keep_prob=TRAIN_KEEP_PROB):
        self.train(None, batch_data, lr, reg_lambda, keep_prob, want_loss=False)

    def wait(self):
        return self.m.wait()

    def eval(self, sess, batch_data, reg_lambda):
        return self.m.eval(sess, batch_data, reg_lambda)   # replicas are identical


class ShardedEmbeddingTrainer(_Base):
    """Row-sharded embedding table + data-parallel dense part (BASELINE.json config 5).

    ``model`` must have been built with feature_size = shard_rows(V_global, world).

    One step, everything on the handle's stream:
      score_prepare_batch -> score_shard_plan (CUDA: positions grouped by owner, counts stay on the device) ->
      ONE all-gather of the [world, world+1] count matrix + its D2H copy (the only host wait of the step, on the copy's
      own event) -> all-to-all ids -> owners: score_shard_presort (side stream) + score_gather_rows -> all-to-all rows
      straight into the staged table -> score_step_begin(staged, mini_keys) (CUDA graph) -> all-reduce dense gradient,
      score_shard_pack_grads, all-to-all gradient rows -> score_step_finish on the owners (presorted keys).
    Asynchronous stepping (train_async) defers a step's score_step_finish into the NEXT call, behind that call's plan
    and count read-back: the device runs the owner-side update while the host waits for the exchange sizes.
    ``ExchangePlan`` above states the same bucketing rule in torch ops (CPU / gloo tests, and the parity test of the
    CUDA plan)."""

    def __init__(self, model, world, rank, group=None, p2p=None):
        super().__init__(model, world, rank, group)
        self._bufs = {}
        # peer-memory exchange (default; SCORE_SHARD_P2P=0 selects the NCCL all-to-alls): the owners store the served rows
        # straight into the requesters' staged tables and the requesters store their gradient rows straight into the
        # owners' buffers over NVLink (kernels of shard.cu, torch symmetric memory for the mappings and the cross-rank
        # barrier) - gather / pack fused with the all-to-all, no NCCL call for the two large exchanges.  Measured on B200
        # (large-vocab step, profiles/r2v / r2w): 1.635 -> 1.462 ms at 2 GPUs, 1.699 -> 1.586 ms at 8; parity as the NCCL path
        self.p2p = ((os.environ.get("SCORE_SHARD_P2P", "1") != "0") if p2p is None else bool(p2p)) and self.world > 1
        self._symm = None
        self._fetch_no = 0
        self._pending = None      # (want, owned, n_recv) of a begun step whose score_step_finish is still to be enqueued
        self.timeline = None      # tools/shard_timeline.py: {phase: [CUDA events]} recorded at the phase boundaries
        self._cm = (C.c_int32 * (self.world * (self.world + 1)))()

    def _buf(self, name, n, dtype):
        """persistent device buffer of at least n elements (grown with slack; the stream is drained before a regrow)"""
        t = self._bufs.get(name)
        if t is None or t.numel() < n:
            if t is not None:
                self.stream.synchronize()
            t = torch.empty(int(n * 1.25) + 1024, dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t[:n]

    def _view(self, ptr, n, dtype):
        ts = "<f4" if dtype == torch.float32 else "<i4"
        return torch.as_tensor(_DevView(ptr, (int(n),), ts), device=self.device)

    def _mark(self, name):
        if self.timeline is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(self.stream)
            self.timeline.setdefault(name, []).append(e)

    def _plan(self, b):
        """-> (send_counts, recv_counts, n_valid, n_recv, plan struct); one host synchronisation."""
        W = self.world
        self._mark("start")
        self.m._check(self.lib.score_prepare_batch(self.h, C.byref(b.struct)))
        plan = _capi.ScoreShardPlan()
        self.m._check(self.lib.score_shard_plan(self.h, W, C.byref(plan)))
        mine = self._view(plan.counts, W + 1, torch.int32)
        if W > 1:
            mat = self._buf("count_mat", W * (W + 1), torch.int32)
            dist.all_gather_into_tensor(mat, mine, group=self.group)
        else:
            mat = mine
        # read-back into pinned memory of the handle + its own event; the previous step's deferred optimizer half goes in
        # between, so the device is busy while the host waits for the sizes
        nn = W * (W + 1)
        self.m._check(self.lib.score_shard_counts_fetch(self.h, mat.data_ptr(), nn))
        self._mark("plan+counts")
        self._flush_finish()
        self._mark("finish(prev)")
        self.m._check(self.lib.score_shard_counts_wait(self.h, self._cm, nn))
        send_counts, recv_counts, self._max_recv, self._max_pos = split_count_matrix(list(self._cm), W, self.rank)
        self._mat = mat
        return send_counts, recv_counts, int(sum(send_counts)), int(sum(recv_counts)), plan

    def _p2p_setup(self, n_positions, d):
        """symmetric buffer [staged table 0 | staged table 1 | gradient rows], the same size on every rank"""
        import torch.distributed._symmetric_memory as symm_mem
        self._n_cap = n_positions + n_positions // 8       # n_positions: the same number on every rank (see _fetch)
        self._owned_cap = 2 * self._n_cap
        self._staged_elems = (self._n_cap + 1) * d
        total = 2 * self._staged_elems + self._owned_cap * d
        self._symm = symm_mem.empty(total, dtype=torch.float32, device=self.device)
        self._symm[:2 * self._staged_elems].zero_()
        grp = self.group if self.group is not None else dist.group.WORLD
        self._symm_hdl = symm_mem.rendezvous(self._symm, grp)
        bases = [int(p) for p in self._symm_hdl.buffer_ptrs]
        self._peer_staged = [(C.c_uint64 * self.world)(*[b + 4 * k * self._staged_elems for b in bases]) for k in (0, 1)]
        self._peer_owned = (C.c_uint64 * self.world)(*[b + 4 * 2 * self._staged_elems for b in bases])
        mine = bases[self.rank]
        self._my_staged = [mine, mine + 4 * self._staged_elems]
        self.m._check(self.lib.score_shard_register_staged(self.h, self._my_staged[0], self._my_staged[1]))
        self.stream.synchronize()

    def _flush_finish(self, want_loss=False):
        """enqueue the optimizer half (score_step_finish) of the step begun last, if it is still pending"""
        if self._pending is None:
            return None
        want, owned, n_recv = self._pending
        self._pending = None
        loss2 = (C.c_float * 2)() if want_loss else None
        self.m._check(self.lib.score_step_finish(self.h, want.data_ptr(), owned.data_ptr(), n_recv, loss2))
        return loss2

    def wait(self):
        """all enqueued steps are complete (asynchronous stepping); returns this rank's last loss share"""
        with torch.cuda.stream(self.stream):
            self._flush_finish()
        return self.m.wait()

    def _fetch(self, b):
        d = self.m.cfg["eb_dim"]
        send_counts, recv_counts, n_valid, n_recv, plan = self._plan(b)
        send_rows = self._view(plan.send_rows, n_valid, torch.int32)
        want = self._buf("want", n_recv, torch.int32)
        if self.world > 1:
            dist.all_to_all_single(want, send_rows, recv_counts, send_counts, group=self.group)
        else:
            want.copy_(send_rows)
        self._mark("a2a ids")
        self._fetch_no += 1
        if self.p2p and (self._symm is None or self._max_pos > self._n_cap or self._max_recv > self._owned_cap):
            # decided from the count matrix alone, which every rank holds: all ranks (re)allocate together
            self.stream.synchronize()
            self._symm = None
            try:
                self._p2p_setup(max(self._max_pos, (self._max_recv + 1) // 2), d)
            except Exception as e:      # no symmetric memory on this system: the NCCL all-to-alls below do the same job
                sys.stderr.write("score_b200: peer-memory exchange unavailable (%s: %s), using NCCL all-to-all\n" % (type(e).__name__, e))
                self.p2p = False
                self._symm = None
        if self.p2p:
            par = self._fetch_no & 1      # alternate the staged tables: a peer may still read the other one in its backward pass
            self.m._check(self.lib.score_shard_serve_push(self.h, want.data_ptr(), n_recv, self._mat.data_ptr(), self.world,
                                                          self.rank, self._peer_staged[par]))
            self._symm_hdl.barrier(channel=0)     # every owner's rows have landed in every staged table
            self._mark("gather")
            self._mark("a2a rows")
            plan.staged = self._my_staged[par]
            return plan, want, (send_counts, recv_counts, n_valid, n_recv)
        served = self._buf("served", n_recv * d, torch.float32).view(n_recv, d)
        self.m._check(self.lib.score_gather_rows(self.h, want.data_ptr(), n_recv, served.data_ptr()))
        self._mark("gather")
        staged = self._view(plan.staged, (plan.n_positions + 1) * d, torch.float32).view(-1, d)
        got = staged[1:1 + n_valid]
        if self.world > 1:
            dist.all_to_all_single(got, served, send_counts, recv_counts, group=self.group)
        else:
            got.copy_(served)
        self._mark("a2a rows")
        return plan, want, (send_counts, recv_counts, n_valid, n_recv)

    def train(self, sess, batch_data, lr, reg_lambda, keep_prob=TRAIN_KEEP_PROB, want_loss=True):
        b = _Batch(batch_data, self.m.cfg)
        gb = b.B * self.world
        d = self.m.cfg["eb_dim"]
        self.m._check(self.lib.score_set_sample_offset(self.h, self.rank * b.B))   # dropout masks of the global batch
        with torch.cuda.stream(self.stream):
            plan, want, (send_counts, recv_counts, n_valid, n_recv) = self._fetch(b)
            self._keep_batch = b
            # the owner-side key list is complete: its sort runs on the side stream under forward / backward
            self.m._check(self.lib.score_shard_presort(self.h, want.data_ptr(), n_recv))
            self.m._check(self.lib.score_step_begin(self.h, None, lr, reg_lambda, keep_prob, gb, 1, plan.staged, plan.mini_keys))
            self._mark("fwd+bwd")
            g = self._dev("dense_grad", torch.float32)
            if self.world > 1:
                dist.all_reduce(g, group=self.group)
            self._mark("allreduce")
            if self.p2p:
                self.m._check(self.lib.score_shard_grad_push(self.h, self._mat.data_ptr(), self.world, self.rank, self._peer_owned))
                self._symm_hdl.barrier(channel=1)     # every requester's gradient rows have landed
                self._mark("pack")
                owned = self._symm[2 * self._staged_elems:2 * self._staged_elems + n_recv * d].view(n_recv, d)
            else:
                self.m._check(self.lib.score_shard_pack_grads(self.h))
                self._mark("pack")
                gsend = self._view(plan.grad_send, n_valid * d, torch.float32).view(n_valid, d)
                owned = self._buf("owned", n_recv * d, torch.float32).view(n_recv, d)
                if self.world > 1:
                    dist.all_to_all_single(owned, gsend, recv_counts, send_counts, group=self.group)
                else:
                    owned.copy_(gsend)
            self._mark("a2a grads")
            self._pending = (want, owned, n_recv)
            if not want_loss:
                return None        # the optimizer half rides behind the next call's plan (or wait())
            return self._global_loss(self._flush_finish(True))

    def train_async(self, batch_data, lr, reg_lambda, keep_prob=TRAIN_KEEP_PROB):
        self.train(None, batch_data, lr, reg_lambda, keep_prob, want_loss=False)

    def eval(self, sess, batch_data, reg_lambda):
        b = _Batch(batch_data, self.m.cfg)
        with torch.cuda.stream(self.stream):
            self._flush_finish()
            plan, want, _ = self._fetch(b)
            self.m._check(self.lib.score_step_begin(self.h, None, 0.0, reg_lambda, 1.0, b.B, 0, plan.staged, plan.mini_keys))
            loss2 = (C.c_float * 2)()
            self.m._check(self.lib.score_step_finish(self.h, None, None, 0, loss2))
            preds = self._dev("y_pred", torch.float32).cpu().numpy().tolist()
        lab = batch_data[6]
        lab = lab.cpu().numpy() if hasattr(lab, "is_cuda") else lab
        import numpy as np
        return preds, np.asarray(lab).astype(np.int32).reshape(-1).tolist(), float(loss2[0])
