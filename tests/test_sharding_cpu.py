"""Host-side logic of the data-parallel path on CPU: world_size-2 gloo process group (no GPU needed).

The sampling path shards independent plans over ranks with no data-path collective; the only exchange is the optional
all-gather of the (B, Ha, A) actions.  These tests cover the row partition, the ragged all-gather, and the fact that
the counter-based noise is keyed by the GLOBAL row (oracle restatement), i.e. shard + row_offset == unsharded."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from latent_diffusion_planning_b200.agent import gather_rows, shard_rows
from oracle import ldp_oracle as O


@pytest.mark.parametrize("n,world", [(1024, 8), (1024, 2), (7, 2), (5, 8), (1, 4)])
def test_shard_rows_partition(n, world):
    spans = [shard_rows(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b and c <= d
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_rows(n, rank, world)
        # each rank "computes" its rows: here the oracle's row-keyed noise, which is what the GPU loops draw
        local = torch.tensor(O.philox_normal_rows(11, 0, 5, lo, hi - lo, 7), dtype=torch.float32)
        full = gather_rows(local, n, world)
        ref = torch.tensor(O.philox_normal_rows(11, 0, 5, 0, n, 7), dtype=torch.float32)
        ok = bool(torch.equal(full, ref))
        t = torch.tensor([float(hi - lo)])
        dist.all_reduce(t)                                  # sum of shard sizes == n on every rank
        q.put((rank, ok, float(t.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7])
def test_gloo_world2_shard_and_gather(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert all(tot == float(n) for _, _, tot in res)
