"""CPU tests of the multi-GPU host logic: conformer range sharding and the single gather of results
(world_size 2, gloo, spawned here; NCCL on the GPU box uses the same code path with CUDA tensors)."""
import os
import socket

import numpy as np
import pytest


def test_shard_ranges_cover_and_balance():
    from mcsce_b200.sharding import shard_range

    for n in (0, 1, 5, 8, 2048, 2049):
        for ws in (1, 2, 3, 8):
            r = [shard_range(n, k, ws) for k in range(ws)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(ws - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, ws, port, n_total, n_atoms, q):
    import torch.distributed as dist

    from mcsce_b200 import sharding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    lo, hi = sharding.my_range(n_total)
    ids = np.arange(lo, hi)
    coords = (ids[:, None, None] * 1000 + np.arange(n_atoms)[None, :, None] + np.arange(3)[None, None, :] * 0.25).astype(np.float32)
    energy = ids.astype(np.float64) * -1.5
    ok = (ids % 3 != 0).astype(np.uint8)
    out = sharding.gather_results(coords, energy, ok, n_total)
    if rank == 0:
        q.put(tuple(np.array(x) for x in out))
    else:
        assert out == (None, None, None)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 8])
def test_gather_two_ranks_gloo(n_total):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_atoms = 5
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, n_atoms, q)) for r in range(2)]
    for p in procs:
        p.start()
    coords, energy, ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ids = np.arange(n_total)
    assert coords.shape == (n_total, n_atoms, 3)
    assert np.array_equal(coords[:, 0, 0], (ids * 1000).astype(np.float32))
    assert np.array_equal(energy, ids * -1.5)
    assert np.array_equal(ok, (ids % 3 != 0).astype(np.uint8))


def _worker_device_result(rank, ws, port, n_total, n_atoms, q):
    import torch
    import torch.distributed as dist

    from mcsce_b200 import sharding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    lo, hi = sharding.my_range(n_total)
    ids = torch.arange(lo, hi)
    coords = (ids[:, None, None] * 10.0 + torch.arange(n_atoms)[None, :, None] + torch.arange(3)[None, None, :] * 0.5).float()
    out = sharding.gather_results(coords, ids.double() * 2.0, (ids % 2).to(torch.uint8), n_total, to_host=False)
    if rank == 0:
        assert all(torch.is_tensor(x) for x in out)             # stays on the gather device, no host copy
        q.put(tuple(x.numpy() for x in out))
    else:
        assert out == (None, None, None)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_keeps_results_on_the_gather_device_when_asked():
    """``to_host=False`` (what bench.py times as the NCCL cross-check of the peer stores): equal shards are gathered straight
    into one buffer, rank 0 gets tensors in global conformer order."""
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_device_result, args=(r, 2, port, 8, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    coords, energy, ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ids = np.arange(8)
    assert np.array_equal(coords[:, 0, 0], (ids * 10.0).astype(np.float32))
    assert np.array_equal(energy, ids * 2.0) and np.array_equal(ok, (ids % 2).astype(np.uint8))


def test_single_process_gather_is_the_identity():
    from mcsce_b200 import sharding

    c, e, k = np.zeros((3, 2, 3), np.float32), np.arange(3.0), np.ones(3, np.uint8)
    out = sharding.gather_results(c, e, k, 3)
    assert out[0] is c and out[1] is e and out[2] is k
    assert sharding.my_range(10) == (0, 10)
