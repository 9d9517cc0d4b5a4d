"""
Data-parallel sharding of a batch of independent designs: one process per GPU, the plan and goal
program replicated, contiguous slices of the batch per rank, and exactly one collective per batched
evaluation — an all-gather of the per-design losses and gradients, or an all-reduce
of their sums in ensemble-mean mode.  There is no exchange inside an evaluation: designs do not couple.

The reference evaluates one design at a time in one process (src/dfdm/optimization.py:89-97); this
module is what lets an ensemble / parameter sweep use 8 x B200.  Works with the ``nccl`` backend on
GPUs and with ``gloo`` on CPU tensors (tests).
"""
import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard", "gather_designs", "gather_to_root", "ensemble_mean", "all_ranks_agree", "ShardedBatch", "PeerGather",
           "bind_to_gpu_cpus"]


def shard_bounds(batch, world_size, rank):
    """Contiguous slice [lo, hi) of ``batch`` designs owned by ``rank``; the first ``batch % world_size`` ranks take one more."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("rank %r outside world of size %r" % (rank, world_size))
    base, extra = divmod(int(batch), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(array, world_size, rank):
    lo, hi = shard_bounds(array.shape[0], world_size, rank)
    return array[lo:hi]


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def gather_designs(local, batch, group=None):
    """
    All-gather per-design rows (``local`` is this rank's [b_local, ...] slice) into [batch, ...] on every
    rank, in design order.  Ragged shards are padded to the largest shard for the collective.
    """
    world, rank = _world(group)
    if world == 1:
        return local
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    most = max(hi - lo for lo, hi in sizes)
    padded = local
    if local.shape[0] != most:
        pad = torch.zeros((most - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    out = torch.empty((world * most,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if all(hi - lo == most for lo, hi in sizes):
        return out
    return torch.cat([out[r * most: r * most + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


def ensemble_mean(loss_local, dq_local, batch, group=None):
    """Mean loss and mean gradient over the whole ensemble: one all-reduce of ``[E + 1]`` doubles."""
    world, _ = _world(group)
    packed = torch.cat([dq_local.sum(dim=0), loss_local.sum().reshape(1)])
    if world > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    packed = packed / float(batch)
    return packed[-1], packed[:-1]


def gather_to_root(local, batch, dst=0, group=None):
    """Per-design rows of every rank on rank ``dst`` only (``None`` elsewhere): what a host-side consumer on one rank
    (a SciPy optimiser over the whole ensemble) needs.  NCCL / gloo ``gather``; equal or ragged shards."""
    world, rank = _world(group)
    if world == 1:
        return local
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    most = max(hi - lo for lo, hi in sizes)
    padded = local.contiguous()
    if local.shape[0] != most:
        pad = torch.zeros((most - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    parts = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, parts, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([parts[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


def all_ranks_agree(ok, device=None, group=None):
    """True on every rank iff ``ok`` is true on every rank (one tiny all-reduce): used before a collective code path is
    taken, so that a failure on some ranks cannot leave the others waiting in a barrier."""
    world, _ = _world(group)
    if world == 1:
        return bool(ok)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device if device is not None else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(int(flag.item()) == 1)


class ShardedBatch:
    """
    Evaluate a global batch with this rank's evaluator and exchange the results.

    ``evaluate(q_local) -> (loss[b_local], dq[b_local, E])`` is the per-rank compute (on the GPU:
    ``Evaluator.loss_and_grad``); the class owns only the partition and the collective.

    ``exchange`` names what the consumer of the results needs (SURVEY.md section 8e allows one collective per
    batched evaluation):

    * ``"allgather"`` (default) - every rank ends up with ``loss[batch]`` and ``dq[batch, E]``: peer stores over NVLink
      when the ranks can map each other's memory (``PeerGather``), NCCL / gloo all-gather otherwise;
    * ``"root"``      - only rank 0 receives them (``(None, None)`` elsewhere): a host-side optimiser on one rank;
    * ``"losses"``    - every rank keeps its own gradients and learns all losses: a sharded optimiser whose state lives
      with the designs needs nothing else (``B`` doubles on the wire);
    * ``"mean"``      - ensemble mean of loss and gradient, one all-reduce of ``E + 1`` doubles (``ensemble``).

    With ``chunks > 1`` the local slice is evaluated in that many pieces and each piece's rows are pushed to the peers on
    a side stream while the next piece computes, so the exchange hides behind the evaluation (peer path only).

    Results of the peer path are views into one of two alternating IPC buffers that the peers overwrite TWO evaluations
    later.  By default they are cloned, so callers may keep them (a previous gradient for BB / L-BFGS, line-search
    trials) exactly like the NCCL path's fresh tensors; ``zero_copy=True`` returns the views themselves for callers
    that consume them before the second next evaluation.
    """

    EXCHANGES = ("allgather", "root", "losses", "mean")

    def __init__(self, evaluate, group=None, peer=True, exchange="allgather", zero_copy=False, chunks=1):
        if exchange not in self.EXCHANGES:
            raise ValueError("exchange must be one of %r" % (self.EXCHANGES,))
        self.evaluate = evaluate
        self.group = group
        self.world_size, self.rank = _world(group)
        self.peer = peer              # gather with peer stores over NVLink when the ranks can (PeerGather.available)
        self.exchange = exchange
        self.zero_copy = zero_copy
        self.chunks = max(1, int(chunks))
        self._gathers = {}
        self.collective = None        # what the last call actually ran (bench.py reports it)
        self.last_local = None        # (loss, dq) of this rank's own designs from the last call

    def local_slice(self, q):
        return shard(q, self.world_size, self.rank)

    # -- peer path ---------------------------------------------------------------------------------------

    def _peer_pair(self, batch, loss, dq):
        """The pair of PeerGather objects for this shape, or None (decided identically on every rank)."""
        if not (self.peer and torch.is_tensor(dq) and dq.is_cuda and PeerGather.available(batch, self.group)):
            return None
        key = (batch, tuple(dq.shape[1:]), dq.dtype)
        if key not in self._gathers:
            pair, err = None, None
            try:
                pair = (PeerGather(batch, (), loss.dtype, loss.device, self.group, collective_setup=False),
                        PeerGather(batch, dq.shape[1:], dq.dtype, dq.device, self.group, collective_setup=False))
                for g in pair:
                    g.exchange_handles()
            except Exception as exc:                        # odd shard sizes, no peer access on this rank, ...
                err = exc
            # every rank reaches this point: the handle exchange above is collective and does not raise before it
            # completes; agree on the outcome so that no rank takes the peer path alone
            if not all_ranks_agree(err is None, device=dq.device, group=self.group):
                if pair is not None:
                    for g in pair:
                        g.close(collective=False)
                pair = None
            self._gathers[key] = pair
        return self._gathers[key]

    def _evaluate_chunked(self, q_local, pair):
        """Evaluate the local slice piece by piece; push every piece's rows to the peers under the next piece's kernels."""
        b = q_local.shape[0]
        k = min(self.chunks, b)
        bounds = [shard_bounds(b, k, c) for c in range(k)]
        g_loss, g_dq = pair
        g_loss.begin()
        g_dq.begin()
        main = torch.cuda.current_stream(g_dq.device)
        side = g_dq.side_stream()
        keep = []
        for lo, hi in bounds:
            loss_c, dq_c = self.evaluate(q_local[lo:hi])
            keep.append((loss_c, dq_c))
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(side):
                side.wait_event(done)
                g_loss.push(loss_c, lo)
                g_dq.push(dq_c, lo)
        main.wait_stream(side)
        loss_all, dq_all = g_loss.end(sync=False), g_dq.end(sync=True)
        self.last_local = keep
        return loss_all, dq_all

    # -- the evaluation ----------------------------------------------------------------------------------

    def loss_and_grad(self, q, gather=True):
        """``q``: the GLOBAL batch [batch, E] (the same on every rank); this rank evaluates its slice."""
        return self.loss_and_grad_local(self.local_slice(q), q.shape[0], gather)

    def loss_and_grad_local(self, q_local, batch, gather=True):
        """The same with only this rank's slice ``q_local`` of a global batch of ``batch`` designs in hand."""
        mode = self.exchange if gather else "none"
        if self.world_size == 1 or mode == "none":
            loss, dq = self.evaluate(q_local)
            self.last_local = [(loss, dq)]
            self.collective = "none"
            return loss, dq
        if mode == "mean":
            loss, dq = self.evaluate(q_local)
            self.last_local = [(loss, dq)]
            self.collective = "all-reduce of E + 1 sums (ensemble mean)"
            return ensemble_mean(loss, dq, batch, self.group)
        if mode == "losses":
            loss, dq = self.evaluate(q_local)
            self.last_local = [(loss, dq)]
            self.collective = "all-gather of the losses only (gradients stay with their designs)"
            return gather_designs(loss, batch, self.group), dq
        if mode == "root":
            loss, dq = self.evaluate(q_local)
            self.last_local = [(loss, dq)]
            self.collective = "gather to rank 0"
            return gather_to_root(loss, batch, 0, self.group), gather_to_root(dq, batch, 0, self.group)
        # all-gather
        probe = None
        if self.peer and q_local.shape[0] > 0:
            # shapes of the results are known without evaluating: loss [b], dq like q
            probe = self._peer_pair(batch, torch.empty(0, dtype=q_local.dtype, device=q_local.device), q_local) \
                if torch.is_tensor(q_local) and q_local.is_cuda else None
        if probe is not None:
            loss_all, dq_all = self._evaluate_chunked(q_local, probe)
            self.collective = "all-gather by peer stores over NVLink (%d piece%s per rank, pushed under the next piece's kernels)" % (
                min(self.chunks, q_local.shape[0]), "" if min(self.chunks, q_local.shape[0]) == 1 else "s")
            if self.zero_copy:
                return loss_all, dq_all
            return loss_all.clone(), dq_all.clone()
        loss, dq = self.evaluate(q_local)
        self.last_local = [(loss, dq)]
        self.collective = "all-gather (%s)" % (dist.get_backend(self.group),)
        return gather_designs(loss, batch, self.group), gather_designs(dq, batch, self.group)

    def ensemble(self, q):
        batch = q.shape[0]
        loss, dq = self.evaluate(self.local_slice(q))
        self.collective = "all-reduce of E + 1 sums (ensemble mean)"
        return ensemble_mean(loss, dq, batch, self.group)

    def close(self):
        """Unmap and free the IPC buffers (collective: call on every rank)."""
        gathers, self._gathers = self._gathers, {}
        for pair in gathers.values():
            if pair is not None:
                for g in pair:
                    g.close()

    def __del__(self):
        # no collective in a finaliser (ranks are finalised at different times): unmap and free locally
        try:
            for pair in getattr(self, "_gathers", {}).values():
                if pair is not None:
                    for g in pair:
                        g.close(collective=False)
        except Exception:
            pass


class _DeviceView:
    """Raw device memory as something ``torch.as_tensor`` can wrap (CUDA array interface, no copy)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


class PeerGather:
    """
    The all-gather of a sharded evaluation done with peer stores over NVLink instead of NCCL.

    Every rank owns ``buffers`` device buffers of ``[batch, *row_shape]`` that all other ranks have mapped through
    CUDA IPC.  ``gather(local)`` launches ONE kernel (``dfdm_peer_push``) that stores this rank's rows into the buffer
    of every rank, the own one included, at the rank's row offset; a one-element all-reduce behind it orders the
    readers: when it completes on a rank, every rank's stores have completed.  The buffers alternate, so a rank may
    still be reading the result of one evaluation while the peers already store the next one.

    Needs equal shards (``batch % world_size == 0``), one process per GPU on one node, and ``torch.distributed``
    initialised.  ``PeerGather.available()`` says whether the peer path can be used; ``gather_designs`` is the NCCL /
    gloo equivalent for every other case.
    """

    _TYPES = {torch.float64: "<f8", torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8"}

    @staticmethod
    def available(batch=None, group=None):
        world, _ = _world(group)
        if world < 2 or not torch.cuda.is_available() or dist.get_backend(group) != "nccl":
            return False
        return batch is None or batch % world == 0

    def __init__(self, batch, row_shape=(), dtype=torch.float64, device=None, group=None, buffers=2, collective_setup=True):
        import ctypes as C
        from dfdm_b200 import _lib
        self._lib, self._C = _lib, C
        self.group = group
        self.world, self.rank = _world(group)
        self._own, self._peers, self._views, self._handles = [], [], [], []
        self._side = None
        if not self.available(batch, group):
            raise RuntimeError("PeerGather needs >= 2 ranks on the nccl backend and a batch divisible by the world size")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shape = (int(batch),) + tuple(int(v) for v in row_shape)
        self.dtype = dtype
        self.rows_local = batch // self.world
        self.row_bytes = int(np.prod(self.shape[1:], dtype=np.int64)) * torch.empty((), dtype=dtype).element_size()
        if self.row_bytes % 8 != 0 or (self.rows_local * self.row_bytes) % 16 != 0:
            raise RuntimeError("PeerGather: rows must be multiples of 8 bytes and a shard a multiple of 16 bytes")
        self._buffers = buffers
        self._turn = 0
        self._cur = None
        self._token = torch.zeros(1, dtype=torch.float32, device=self.device)
        if collective_setup:
            self.exchange_handles()

    def exchange_handles(self):
        """Allocate the own buffers and map every peer's (collective).  A local failure is carried THROUGH the handle
        exchange (the rank publishes ``None``) and raised on every rank after it, so no rank is left waiting."""
        C, _lib = self._C, self._lib
        total = self.shape[0] * self.row_bytes
        failure = None
        with torch.cuda.device(self.device):
            for _ in range(self._buffers):
                ptr = C.c_void_p()
                handle = C.create_string_buffer(64)
                raw = None
                if failure is None:
                    try:
                        _lib.check(_lib.lib.dfdm_peer_alloc(total, C.byref(ptr), handle))
                        raw = handle.raw
                        self._own.append(ptr.value)
                    except Exception as exc:
                        failure = exc
                handles = [None] * self.world
                dist.all_gather_object(handles, raw, group=self.group)
                if failure is None and any(h is None for h in handles):
                    failure = RuntimeError("PeerGather: a peer could not allocate its buffer")
                bases = []
                if failure is None:
                    for r, h in enumerate(handles):
                        if r == self.rank:
                            bases.append(ptr.value)
                            continue
                        p = C.c_void_p()
                        try:
                            _lib.check(_lib.lib.dfdm_peer_open(h, C.byref(p)))
                            bases.append(p.value)
                        except Exception as exc:
                            failure = exc
                            break
                if failure is None:
                    self._peers.append(bases)
                    self._views.append(torch.as_tensor(_DeviceView(ptr.value, self.shape, self._TYPES[self.dtype]), device=self.device))
                else:
                    self._peers.append([b for b in bases])
        ok = all_ranks_agree(failure is None, device=self.device, group=self.group)
        if not ok:
            self.close(collective=False)
            raise RuntimeError("PeerGather: peer mapping failed on at least one rank (%s)" % (failure if failure else "a peer",))

    def side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    # -- one gather = begin, any number of pushes of consecutive local rows, end --------------------------

    def begin(self):
        """Take the next buffer (they alternate: a rank may still read one result while its peers store the next)."""
        self._cur = self._turn
        self._turn = (self._turn + 1) % len(self._own)
        return self._cur

    def push(self, rows, first_local_row=0, only=None):
        """Store ``rows`` (local rows ``first_local_row ...``) into every rank's buffer (or into rank ``only``'s) on the
        current stream: ONE kernel, plain stores to peer addresses."""
        C = self._C
        rows = rows.contiguous()
        if tuple(rows.shape[1:]) != self.shape[1:] or rows.dtype != self.dtype or first_local_row + rows.shape[0] > self.rows_local:
            raise ValueError("PeerGather.push: rows %r of %s do not fit a shard of %d x %r" % (tuple(rows.shape), rows.dtype, self.rows_local, self.shape[1:]))
        nbytes = rows.shape[0] * self.row_bytes
        offset = (self.rank * self.rows_local + first_local_row) * self.row_bytes
        if nbytes == 0:
            return
        if offset % 16 != 0 or rows.data_ptr() % 16 != 0:
            raise ValueError("PeerGather.push: pieces must start on 16-byte boundaries")
        targets = self._peers[self._cur] if only is None else [self._peers[self._cur][only]]
        bases = (C.c_void_p * len(targets))(*targets)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        self._lib.check(self._lib.lib.dfdm_peer_push(C.c_void_p(rows.data_ptr()), nbytes, offset, bases, len(targets), stream))

    def end(self, sync=True):
        view = self._views[self._cur]
        if sync:
            self.fence()
        return view

    def gather(self, local, sync=True):
        """Store ``local`` (this rank's ``[batch / world, ...]`` rows) into every rank's buffer; returns the own, full one
        (a view that the peers overwrite two gathers later: clone it to keep it)."""
        if tuple(local.shape) != (self.rows_local,) + self.shape[1:] or local.dtype != self.dtype:
            raise ValueError("PeerGather.gather: expected %r of %s" % ((self.rows_local,) + self.shape[1:], self.dtype))
        self.begin()
        self.push(local, 0)
        return self.end(sync)

    def fence(self):
        """Orders this rank's readers after every rank's stores (a one-element all-reduce on the current stream)."""
        dist.all_reduce(self._token, group=self.group)

    def close(self, collective=True):
        C = self._C
        for bases in self._peers:
            for r, p in enumerate(bases):
                if r != self.rank and p:
                    self._lib.lib.dfdm_peer_close(C.c_void_p(p))
        if collective and dist.is_available() and dist.is_initialized():
            dist.barrier(group=self.group)        # nobody frees a buffer a peer still has mapped
        for own in self._own:
            self._lib.lib.dfdm_peer_free(C.c_void_p(own))
        self._peers, self._own, self._views = [], [], []


def bind_to_gpu_cpus(device_index):
    """
    Pin this process to the CPU cores next to its GPU (NVML's affinity mask for the device, cut to the cores the
    process may use).  With one process per GPU on a multi-socket host this keeps pinned host buffers on the GPU's own
    NUMA node, which is what the host-buffer entry points copy from and to.  Returns the cores, or None when NVML,
    the device or the platform does not allow it (nothing is changed then).
    """
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        near = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = near & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return sorted(allowed)
    except Exception:
        return None
