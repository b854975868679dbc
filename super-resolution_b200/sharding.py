"""Multi-GPU frame sharding of the MAP objective, one process per GPU.

The data term is a sum over LR frames (objective_data_term.cpp:104-114), so rank r of G holds a
contiguous block of the frames and their observations; the estimate x and the IRLS weights are
replicated; the regularization term is split by HR row bands so that the sum over ranks is the
full objective.  ONE allreduce(sum) over C*P + 1 doubles (gradient, cost in the last slot) per
evaluation gives gradient and cost on every rank.

The allreduce is pipelined against the computation: the gradient is produced in contiguous
"units" (srb_eval_units_dev), and the slice of the units already computed is reduced over
NVLink (NCCL, its own stream) while the tile kernel works on the next ones.

`ShardedObjective` only needs an evaluator with the four methods of `EngineEvaluator`; the CPU
tests drive it with an oracle-backed evaluator over gloo (world_size 2).
"""
import numpy as np


def frame_shard(num_frames, rank, world):
    """Frame indices owned by `rank`: contiguous blocks.  (Any partition of the frames is valid, the
    data term is a plain sum over them; blocks keep a capture sequence that cycles through the
    sub-pixel phases -- one frame per phase and block -- balanced on every rank, which is what the
    tile kernel's table-driven residual pass wants.)"""
    base, extra = divmod(int(num_frames), int(world))
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def row_band(H, rank, world):
    """HR row band [r0, r1) whose regularization term `rank` evaluates."""
    band = (H + world - 1) // world
    return min(H, rank * band), min(H, (rank + 1) * band)


def chunk_bounds(num_units, num_chunks):
    """Splits units [0, num_units) into at most num_chunks contiguous, near-equal chunks."""
    num_chunks = max(1, min(int(num_chunks), int(num_units)))
    edges = [(i * num_units) // num_chunks for i in range(num_chunks + 1)]
    return [(edges[i], edges[i + 1]) for i in range(num_chunks) if edges[i + 1] > edges[i]]


import contextlib


def engine_stream(evaluator, tensor):
    """Context in which torch's CURRENT stream is the engine's own CUDA stream.  The library launches its kernels on
    the context's private stream while NCCL orders its collectives against torch's current stream: a collective on
    a gradient slice may only start once the kernels that write the slice are done, so both must be the same
    stream.  A no-op for CPU tensors and for evaluators without an engine (the gloo tests)."""
    eng = getattr(evaluator, "e", evaluator)
    if not getattr(tensor, "is_cuda", False) or not hasattr(eng, "stream_handle"):
        return contextlib.nullcontext()
    import torch
    return torch.cuda.stream(torch.cuda.ExternalStream(eng.stream_handle(), device=tensor.device))


def stencil_halo_rows(psf_size, reg_kind=-1, btv_range=3):
    """HR rows of x around a gradient row that the fused kernels read: the PSF twice (forward and adjoint pass) plus
    one row for TV / 3-D TV or R rows for BTV (csrc/srb_multi.cuh: multi_halo_rows)."""
    reg = btv_range if reg_kind == 2 else 1
    return 2 * (int(psf_size) // 2) + max(int(reg), 1)


def unit_band(num_units, rank, world):
    """Contiguous units [u0, u1) of `rank` in the row-band partition (units are (channel, tile row) in memory
    order, so a rank's gradient rows are one contiguous element range)."""
    return (rank * num_units) // world, ((rank + 1) * num_units) // world


class RowBandObjective:
    """Row-band partition of the MAP objective over the ranks (the alternative to frame sharding): every rank
    holds EVERY frame and evaluates the whole objective on its contiguous band of (channel, tile row) units.  The
    gradient band a rank produces is final -- there is no cross-rank sum of the gradient; what the ranks exchange
    per evaluation is the halo of the estimate (the few rows next to a band that the PSF / regularizer stencils
    of the neighbouring band read: `halo_rows` rows of W doubles to each neighbour) and one scalar (the cost).

    In a distributed solver every rank updates its own band of x; `evaluate` therefore first refreshes the halo
    rows of this rank's replica from the neighbours' bands, then evaluates its units.

    evaluator protocol: num_units(), unit_range(u0, u1) -> element range, eval_unit_range(x, g, u0, u1, cost)
    with `cost` a 1-element tensor on x's device.
    """

    def __init__(self, evaluator, n, width, halo_rows, dist=None, group=None):
        self.ev, self.n, self.dist, self.group = evaluator, int(n), dist, group
        self.world = 1 if dist is None or not dist.is_initialized() else dist.get_world_size(group)
        self.rank = 0 if self.world == 1 else dist.get_rank(group)
        nu = evaluator.total_units() if hasattr(evaluator, "total_units") else evaluator.num_units()
        rng = evaluator.row_unit_range if hasattr(evaluator, "row_unit_range") else evaluator.unit_range
        self.u0, self.u1 = unit_band(nu, self.rank, self.world)
        self.begin, self.end = rng(self.u0, self.u1)
        self.halo = int(halo_rows) * int(width)            # elements exchanged with each neighbour
        # neighbours with a non-empty band (ranks beyond the number of units hold nothing)
        bands = [unit_band(nu, r, self.world) for r in range(self.world)]
        self.prev = next((r for r in range(self.rank - 1, -1, -1) if bands[r][1] > bands[r][0]), None)
        self.next = next((r for r in range(self.rank + 1, self.world) if bands[r][1] > bands[r][0]), None)
        self.empty = self.u1 <= self.u0
        for r in range(self.world):
            b, e = rng(*bands[r])
            if bands[r][1] > bands[r][0] and e - b < self.halo and self.world > 1:
                raise ValueError("a rank's row band is thinner than the stencil halo: use fewer ranks")

    def exchange_halo(self, x):
        """x[begin - halo, begin) <- previous rank's last rows; x[end, end + halo) <- next rank's first rows."""
        if self.world == 1 or self.empty:
            return
        d = self.dist
        ops = []
        h = min(self.halo, self.end - self.begin)
        if self.prev is not None:
            ops.append(d.P2POp(d.isend, x[self.begin:self.begin + h], self.prev, self.group))
            ops.append(d.P2POp(d.irecv, x[max(self.begin - self.halo, 0):self.begin], self.prev, self.group))
        if self.next is not None:
            ops.append(d.P2POp(d.isend, x[self.end - h:self.end], self.next, self.group))
            ops.append(d.P2POp(d.irecv, x[self.end:min(self.end + self.halo, self.n)], self.next, self.group))
        for w in d.batch_isend_irecv(ops):
            w.wait()

    def evaluate(self, x, g, cost):
        """g[begin:end] <- this rank's final gradient band; cost[0] <- the whole objective's cost (all ranks).
        Runs on the engine's stream (engine_stream): halo, kernels and the scalar sum are ordered on one stream."""
        with engine_stream(self.ev, x):
            self.exchange_halo(x)
            self.ev.eval_unit_range(x, g, self.u0, self.u1, cost)
            if self.world > 1:
                self.dist.all_reduce(cost, group=self.group)


class EngineEvaluator:
    """Adapter of one `Engine` (one rank's srb_ctx) to the evaluator protocol."""

    def __init__(self, engine):
        self.e = engine

    def num_units(self):
        return self.e.num_units()[0]

    def unit_range(self, u0, u1):
        return self.e.unit_range(u0, u1)

    def total_units(self):
        return self.e.total_units()

    def row_unit_range(self, u0, u1):
        """Element range of (channel, 32-row tile) units [u0, u1): valid whether or not the model pipelines."""
        tr = (self.e.H + 31) // 32
        P = self.e.H * self.e.W

        def first(u):
            ch, t = divmod(u, tr)
            return ch * P + min(t * 32, self.e.H) * self.e.W
        return first(u0), first(u1)

    def eval_units(self, x, gc, u0, u1):
        self.e.eval_units_dev(x, gc, u0, u1)

    def eval_finish(self, x, gc):
        self.e.eval_finish_dev(x, gc)

    def eval_unit_range(self, x, g, u0, u1, cost):
        self.e.eval_unit_range_dev(x, g, u0, u1, cost)


class ShardedObjective:
    """cost + gradient of the full objective from per-rank partial evaluations.

    evaluate(x, gc): x is the replicated estimate, gc a buffer of n + 1 doubles; on return (after
    `wait()`, or immediately for synchronous backends) gc[:n] is the full gradient and gc[n] the
    full cost on every rank.  With a CUDA engine behind the evaluator the whole evaluation -- kernels and
    collectives -- is issued with the engine's own stream as torch's current stream (engine_stream), so that every
    allreduce is ordered behind the kernels that wrote its slice; results are valid on that stream.
    """

    def __init__(self, evaluator, n, dist=None, group=None, num_chunks=4):
        self.ev = evaluator
        self.n = int(n)
        self.dist = dist            # torch.distributed (None or world size 1: no collective)
        self.group = group
        self.num_chunks = num_chunks
        self._pending = []
        self._agreed_units = None

    def _world(self):
        if self.dist is None or not self.dist.is_initialized():
            return 1
        return self.dist.get_world_size(self.group)

    def _units(self, gc):
        """Number of units every rank cuts its gradient into.  Ranks must issue the SAME sequence of
        collectives, but whether a rank can pipeline depends on its own frames (a shift beyond the
        PSF half width puts border-band samples on that rank only): agree once on the minimum;
        1 means "no pipelining anywhere"."""
        mine = self.ev.num_units()
        if self._world() == 1:
            return mine
        if self._agreed_units is None:
            import torch
            t = torch.tensor([mine, -mine], dtype=torch.int64, device=gc.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
            lo, hi = int(t[0]), -int(t[1])
            self._agreed_units = lo if lo == hi else 1
        return self._agreed_units if self._agreed_units == mine else 1

    def evaluate(self, x, gc):
        with engine_stream(self.ev, gc):
            return self._evaluate(x, gc)

    def _evaluate(self, x, gc):
        world = self._world()
        units = self._units(gc)
        chunks = chunk_bounds(units, self.num_chunks if world > 1 else 1)
        if units == 1 and self.ev.num_units() != 1:
            chunks = [(0, self.ev.num_units())]   # this rank could pipeline, another cannot
        self._pending = []
        for i, (u0, u1) in enumerate(chunks):
            self.ev.eval_units(x, gc, u0, u1)
            last = i == len(chunks) - 1
            if last:
                self.ev.eval_finish(x, gc)
            if world > 1:
                b, e = self.ev.unit_range(u0, u1)
                if last:
                    e = self.n + 1      # the cost slot rides with the last slice
                work = self.dist.all_reduce(gc[b:e], group=self.group, async_op=True)
                self._pending.append(work)
        return self

    def wait(self):
        for w in self._pending:
            w.wait()
        self._pending = []


class _DevArray:
    """__cuda_array_interface__ view of a raw device pointer (float64, 1-D)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


class PeerObjective:
    """cost + gradient of the full objective over NVLink peer memory (srb_peer_* in srb200.h).
    Every rank owns a contiguous band of the gradient.  The tile kernel evaluates the partial
    gradient band by band, other ranks' bands first; each finished band is pushed by a copy engine
    into the owner's slot array (CUDA-IPC peer mappings over NVLink / NVSwitch) while the SMs compute
    the next band.  The owners then sum their band in fixed rank order and store the result into every
    rank's gradient buffer (the all-gather half).  The barriers between the phases are flags in peer
    memory (bounded spin, no host synchronisation); SRB_PEER_NCCL_BARRIER=1 adds one-element NCCL
    allreduces as well.  Construction is all-or-nothing across the ranks: it raises on EVERY rank when
    the model does not qualify on some rank (border band, non-fused regularizer) or a peer buffer
    cannot be mapped -- use ShardedObjective then.

    evaluate(x): on return (stream-ordered) `self.out[:n]` holds the full gradient and
    `self.out[n]` the full cost on every rank."""

    def __init__(self, engine, n, dist, pkg, group=None):
        import torch
        self.e, self.n, self.dist, self.group, self.pkg = engine, int(n), dist, group, pkg
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        # whether a rank qualifies depends on its own frames (border bands): agree first, so that
        # either every rank sets the peer path up or every rank raises
        try:
            slots_bytes, out_bytes = engine.peer_sizes(self.world)
            err = None
        except Exception as exc:   # SrbError(SRB_ERR_STATE)
            slots_bytes = out_bytes = 0
            err = exc
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok[0]) == 0:
            raise err if err is not None else pkg.SrbError(4, "another rank cannot take the peer path")
        self._slots = pkg.dev_alloc(slots_bytes)
        self._out = pkg.dev_alloc(out_bytes)
        mine = (pkg.ipc_export(self._slots), pkg.ipc_export(self._out))
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        self._opened = []
        slot_ptrs, out_ptrs = [], []
        try:
            for o, (hs, ho) in enumerate(handles):
                if o == self.rank:
                    slot_ptrs.append(self._slots)
                    out_ptrs.append(self._out)
                else:
                    ps = pkg.ipc_open(hs)
                    self._opened.append(ps)
                    po = pkg.ipc_open(ho)
                    self._opened.append(po)
                    slot_ptrs.append(ps)
                    out_ptrs.append(po)
            engine.peer_setup(self.rank, self.world, slot_ptrs, out_ptrs)
            err = None
        except Exception as exc:   # e.g. no peer access between two of the GPUs
            err = exc
        # again all or nothing: a rank that could not map a peer must not leave the others waiting
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok[0]) == 0:
            for p in self._opened:
                pkg.ipc_close(p)
            self._opened = []
            dist.barrier(group=group)
            pkg.dev_free(self._slots)
            pkg.dev_free(self._out)
            raise err if err is not None else pkg.SrbError(2, "another rank could not map the peer buffers")
        self.slot_ptrs, self.out_ptrs = slot_ptrs, out_ptrs
        self.out = torch.as_tensor(_DevArray(self._out, self.n + 1 + 3 * self.world), device="cuda")
        import os
        self._nccl_barrier = os.environ.get("SRB_PEER_NCCL_BARRIER", "0") == "1"
        self._flag = torch.zeros(1, dtype=torch.float32, device="cuda")
        dist.barrier(group=group)

    def _barrier(self):
        # stream-ordered device barrier: nobody passes until every rank's preceding kernels are done
        self.dist.all_reduce(self._flag, group=self.group)

    def evaluate(self, x):
        with engine_stream(self.e, x):   # (only the optional NCCL barriers care: the peer kernels are the engine's)
            self.e.peer_scatter_dev(x)
            if self._nccl_barrier:
                self._barrier()
            self.e.peer_gather_dev()
            if self._nccl_barrier:
                self._barrier()
        return self

    def wait(self):
        return self

    def close(self):
        self.dist.barrier(group=self.group)
        for p in self._opened:
            self.pkg.ipc_close(p)
        self._opened = []
        self.out = None
        self.dist.barrier(group=self.group)
        self.pkg.dev_free(self._slots)
        self.pkg.dev_free(self._out)


def reduce_reference(partials):
    """What the collective computes: elementwise sum of the ranks' (gradient, cost) buffers."""
    return np.sum(np.stack(partials, axis=0), axis=0)
