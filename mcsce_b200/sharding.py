"""Conformer sharding across GPUs (one process per GPU, ``torchrun``).

The reference's only parallelism is a ``multiprocessing.Pool`` over independent conformers
(``side_chain_builder.py:260-274, 328-347``).  Here a call's conformers ``[0, n)`` are split into
contiguous ranges, one per rank; RNG streams are keyed by the global conformer id, so the result does
not depend on the number of ranks (bit for bit while every rank holds enough conformers to fill the GPU
without splitting the environment, about 150 x trial tiles; below that the environment split regroups
float32 partial sums and conformers agree to rounding).  There is no collective on the growth path.
Results reach rank 0 in one of two ways:

* :class:`PeerResults` (the product path on one NVSwitch node): rank 0 owns the ensemble's result buffer, every
  other rank maps it through CUDA IPC and hands its slice to ``mcsce_grow`` as the output pointers, so the result
  kernels store straight into rank 0's HBM over NVLink - the "gather" is the kernels' own store stream;
* :func:`gather_results`: one ``torch.distributed`` gather per array (NCCL for CUDA tensors, gloo for the CPU
  tests) - used when the peer mapping is not available (several nodes, no IPC) and as the checker of the above.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def shard_range(n_total, rank, world_size):
    """Contiguous, balanced split of ``range(n_total)``: sizes differ by at most one."""
    base, extra = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def my_range(n_total):
    rank, ws = world()
    return shard_range(n_total, rank, ws)


def gather_results(coords, energy, ok, n_total, to_host=True):
    """Gather per-rank result blocks to rank 0 in global conformer order.

    ``coords (c, N, 3) f32``, ``energy (c,) f64``, ``ok (c,) u8`` as NumPy arrays or torch tensors; returns the
    concatenated arrays on rank 0 (NumPy, or torch tensors on the gather device with ``to_host=False``) and
    ``(None, None, None)`` on the other ranks."""
    rank, ws = world()
    if ws == 1:
        if not to_host:
            return coords, energy, ok
        to_np = lambda x: x if isinstance(x, np.ndarray) else x.cpu().numpy()  # noqa: E731
        return to_np(coords), to_np(energy), to_np(ok)
    import torch
    import torch.distributed as dist

    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    as_t = lambda x, dt: (torch.as_tensor(x).to(device=dev, dtype=dt)).contiguous()  # noqa: E731
    c, e, k = as_t(coords, torch.float32), as_t(energy, torch.float64), as_t(ok, torch.uint8)
    n_atoms = c.shape[1]
    sizes = [shard_range(n_total, r, ws)[1] - shard_range(n_total, r, ws)[0] for r in range(ws)]
    cap = max(sizes)
    even = min(sizes) == cap           # equal blocks: gather straight into one buffer, no padding copies

    def pad(t, shape):
        if even:
            return t
        out = torch.zeros(shape, dtype=t.dtype, device=dev)
        out[: t.shape[0]] = t
        return out

    bufs = [pad(c, (cap, n_atoms, 3)), pad(e, (cap,)), pad(k, (cap,))]
    gathered = []
    for b in bufs:                                   # three gathers in total, once per ensemble call
        lst = None
        if rank == 0:
            whole = torch.empty((ws * cap,) + tuple(b.shape[1:]), dtype=b.dtype, device=dev)
            lst = list(whole.split(cap))
        dist.gather(b, lst, dst=0)
        gathered.append((whole, lst) if rank == 0 else None)
    if rank != 0:
        return None, None, None
    out = []
    for whole, lst in gathered:
        t = whole if even else torch.cat([x[:n] for x, n in zip(lst, sizes)])
        out.append(t.cpu().numpy() if to_host else t)
    return tuple(out)


class _DeviceBlock:
    """Raw device memory as a ``__cuda_array_interface__`` provider, so torch can view it without owning it."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class PeerResults:
    """One result buffer for the whole ensemble in rank 0's HBM, mapped by every rank of the node (CUDA IPC).

    ``slice_ptrs(lo, hi)`` gives the (coords, energy, ok) device addresses of conformers ``[lo, hi)`` for
    ``GrowthEngine.grow(out_ptrs=...)``; after ``fence()`` (a barrier that orders the peers' stores before rank 0's
    reads) rank 0 reads the complete arrays with ``tensors()``.  Collective: construct on every rank."""

    def __init__(self, n_total, n_atoms, device_index):
        import torch
        import torch.distributed as dist

        from mcsce_b200._lib import check, load_library

        self.lib, self.check = load_library(), check
        self.rank, self.ws = world()
        self.n_total, self.n_atoms, self.device_index = int(n_total), int(n_atoms), int(device_index)
        al = lambda v: (v + 255) // 256 * 256  # noqa: E731
        self.off_energy = al(self.n_total * self.n_atoms * 12)
        self.off_ok = self.off_energy + al(self.n_total * 8)
        self.bytes = self.off_ok + al(self.n_total)
        self.base = None
        self.owner = self.rank == 0
        handle = C.create_string_buffer(64)
        if self.owner:
            p = C.c_void_p()
            check(self.lib.mcsce_peer_alloc(self.device_index, self.bytes, C.byref(p), handle))
            self.base = int(p.value)
        if self.ws > 1:
            box = [bytes(handle.raw) if self.owner else None]
            dist.broadcast_object_list(box, src=0)
            if not self.owner:
                p = C.c_void_p()
                check(self.lib.mcsce_peer_open(self.device_index, C.create_string_buffer(box[0], 64), C.byref(p)))
                self.base = int(p.value)
        self._torch = torch

    def slice_ptrs(self, lo, hi):
        assert 0 <= lo <= hi <= self.n_total
        return (self.base + lo * self.n_atoms * 12, self.base + self.off_energy + lo * 8, self.base + self.off_ok + lo)

    def fence(self):
        """Every rank's result kernels have finished and their stores are visible to rank 0."""
        self._torch.cuda.synchronize()
        if self.ws > 1:
            import torch.distributed as dist
            dist.barrier()

    def tensors(self):
        """(coords (n, N, 3) f32, energy (n,) f64, ok (n,) u8) views of the buffer; rank 0 only."""
        assert self.owner, "only rank 0 owns the result buffer"
        torch, n = self._torch, self.n_total
        dev = torch.device("cuda", self.device_index)
        c = torch.as_tensor(_DeviceBlock(self.base, (n, self.n_atoms, 3), "<f4"), device=dev)
        e = torch.as_tensor(_DeviceBlock(self.base + self.off_energy, (n,), "<f8"), device=dev)
        k = torch.as_tensor(_DeviceBlock(self.base + self.off_ok, (n,), "|u1"), device=dev)
        return c, e, k

    def close(self):
        if self.base is None:
            return
        if self.owner:
            if self.ws > 1:
                import torch.distributed as dist
                dist.barrier()                      # peers unmap before the owner frees
            self.lib.mcsce_peer_free(self.device_index, C.c_void_p(self.base))
        else:
            self.lib.mcsce_peer_close(self.device_index, C.c_void_p(self.base))
            import torch.distributed as dist
            dist.barrier()
        self.base = None
