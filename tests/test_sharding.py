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
