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

__all__ = ["shard_bounds", "shard", "gather_designs", "ensemble_mean", "ShardedBatch", "PeerGather", "bind_to_gpu_cpus"]


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


class ShardedBatch:
    """
    Evaluate a global batch with this rank's evaluator and gather the results.

    ``evaluate(q_local) -> (loss[b_local], dq[b_local, E])`` is the per-rank compute (on the GPU:
    ``Evaluator.loss_and_grad``); the class owns only the partition and the collective.
    """

    def __init__(self, evaluate, group=None, peer=True):
        self.evaluate = evaluate
        self.group = group
        self.world_size, self.rank = _world(group)
        self.peer = peer              # gather with peer stores over NVLink when the ranks can (PeerGather.available)
        self._gathers = {}

    def local_slice(self, q):
        return shard(q, self.world_size, self.rank)

    def loss_and_grad(self, q, gather=True):
        batch = q.shape[0]
        loss, dq = self.evaluate(self.local_slice(q))
        if not gather:
            return loss, dq
        if self.peer and torch.is_tensor(dq) and dq.is_cuda and PeerGather.available(batch, self.group):
            key = (batch, tuple(dq.shape[1:]), dq.dtype)
            if key not in self._gathers:
                try:
                    self._gathers[key] = (PeerGather(batch, (), loss.dtype, loss.device, self.group),
                                          PeerGather(batch, dq.shape[1:], dq.dtype, dq.device, self.group))
                except RuntimeError:                    # odd shard sizes, no peer access: every rank takes the NCCL path
                    self._gathers[key] = None
            if self._gathers[key] is not None:
                g_loss, g_dq = self._gathers[key]
                loss_all, dq_all = g_loss.gather(loss, sync=False), g_dq.gather(dq, sync=False)
                g_dq.fence()
                return loss_all, dq_all
        return gather_designs(loss, batch, self.group), gather_designs(dq, batch, self.group)

    def ensemble(self, q):
        batch = q.shape[0]
        loss, dq = self.evaluate(self.local_slice(q))
        return ensemble_mean(loss, dq, batch, self.group)


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

    def __init__(self, batch, row_shape=(), dtype=torch.float64, device=None, group=None, buffers=2):
        import ctypes as C
        from dfdm_b200 import _lib
        self._lib, self._C = _lib, C
        self.group = group
        self.world, self.rank = _world(group)
        if not self.available(batch, group):
            raise RuntimeError("PeerGather needs >= 2 ranks on the nccl backend and a batch divisible by the world size")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shape = (int(batch),) + tuple(int(v) for v in row_shape)
        self.dtype = dtype
        self.rows_local = batch // self.world
        self.row_bytes = int(np.prod(self.shape[1:], dtype=np.int64)) * torch.empty((), dtype=dtype).element_size()
        if (self.rows_local * self.row_bytes) % 16 != 0:
            raise RuntimeError("PeerGather: a shard must be a multiple of 16 bytes")
        total = self.shape[0] * self.row_bytes
        self._own, self._peers, self._views = [], [], []
        with torch.cuda.device(self.device):
            for _ in range(buffers):
                ptr = C.c_void_p()
                handle = C.create_string_buffer(64)
                _lib.check(_lib.lib.dfdm_peer_alloc(total, C.byref(ptr), handle))
                handles = [None] * self.world
                dist.all_gather_object(handles, handle.raw, group=group)
                bases = []
                for r, h in enumerate(handles):
                    if r == self.rank:
                        bases.append(ptr.value)
                    else:
                        p = C.c_void_p()
                        _lib.check(_lib.lib.dfdm_peer_open(h, C.byref(p)))
                        bases.append(p.value)
                self._own.append(ptr.value)
                self._peers.append(bases)
                self._views.append(torch.as_tensor(_DeviceView(ptr.value, self.shape, self._TYPES[dtype]), device=self.device))
        self._turn = 0
        self._token = torch.zeros(1, dtype=torch.float32, device=self.device)
        dist.barrier(group=group)

    def gather(self, local, sync=True):
        """Store ``local`` (this rank's ``[batch / world, ...]`` rows) into every rank's buffer; returns the own, full one."""
        C = self._C
        local = local.contiguous()
        if tuple(local.shape) != (self.rows_local,) + self.shape[1:] or local.dtype != self.dtype:
            raise ValueError("PeerGather.gather: expected %r of %s" % ((self.rows_local,) + self.shape[1:], self.dtype))
        k = self._turn
        self._turn = (k + 1) % len(self._own)
        bases = (C.c_void_p * self.world)(*self._peers[k])
        nbytes = self.rows_local * self.row_bytes
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        self._lib.check(self._lib.lib.dfdm_peer_push(C.c_void_p(local.data_ptr()), nbytes, self.rank * nbytes, bases, self.world, stream))
        if sync:
            self.fence()
        return self._views[k]

    def fence(self):
        """Orders this rank's readers after every rank's stores (a one-element all-reduce on the current stream)."""
        dist.all_reduce(self._token, group=self.group)

    def close(self):
        for bases, own in zip(self._peers, self._own):
            for r, p in enumerate(bases):
                if r != self.rank:
                    self._lib.lib.dfdm_peer_close(self._C.c_void_p(p))
        if dist.is_initialized():
            dist.barrier(group=self.group)
        for own in self._own:
            self._lib.lib.dfdm_peer_free(self._C.c_void_p(own))
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
