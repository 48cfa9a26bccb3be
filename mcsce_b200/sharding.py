"""Conformer sharding across GPUs (one process per GPU, ``torchrun``).

The reference's only parallelism is a ``multiprocessing.Pool`` over independent conformers
(``side_chain_builder.py:260-274, 328-347``).  Here a call's conformers ``[0, n)`` are split into
contiguous ranges, one per rank; RNG streams are keyed by the global conformer id, so the result does
not depend on the number of ranks (bit for bit while every rank holds enough conformers to fill the GPU
without splitting the environment, about 150 x trial tiles; below that the environment split regroups
float32 partial sums and conformers agree to rounding).  There is no collective on the growth path; the final coordinates,
energies and success flags are gathered to rank 0 once (NCCL over NVLink when the ranks hold CUDA
tensors, gloo for the CPU tests).
"""
from __future__ import annotations

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


def gather_results(coords, energy, ok, n_total):
    """Gather per-rank result blocks to rank 0 in global conformer order.

    ``coords (c, N, 3) f32``, ``energy (c,) f64``, ``ok (c,) u8`` as NumPy arrays or torch tensors; returns the
    concatenated NumPy arrays on rank 0 and ``(None, None, None)`` on the other ranks."""
    rank, ws = world()
    if ws == 1:
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

    def pad(t, shape):
        out = torch.zeros(shape, dtype=t.dtype, device=dev)
        out[: t.shape[0]] = t
        return out

    bufs = [pad(c, (cap, n_atoms, 3)), pad(e, (cap,)), pad(k, (cap,))]
    gathered = []
    for b in bufs:                                   # three gathers in total, once per ensemble call
        lst = [torch.empty_like(b) for _ in range(ws)] if rank == 0 else None
        dist.gather(b, lst, dst=0)
        gathered.append(lst)
    if rank != 0:
        return None, None, None
    cat = lambda lst: torch.cat([t[:n] for t, n in zip(lst, sizes)]).cpu().numpy()  # noqa: E731
    return cat(gathered[0]), cat(gathered[1]), cat(gathered[2])
