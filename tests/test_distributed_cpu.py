"""world_size-2 gloo tests of the multi-process host logic (chain sharding, diagnostics gather)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pymc4_b200 import diagnostics, parallel


def test_shard_chains_partitions_exactly():
    for total, world in ((4096, 1), (4096, 8), (10, 4), (7, 8), (65536, 8)):
        blocks = [parallel.shard_chains(total, world, r) for r in range(world)]
        assert sum(c for c, _ in blocks) == total
        pos = 0
        for count, off in blocks:
            assert off == pos
            pos += count
    with pytest.raises(ValueError):
        parallel.shard_chains(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total, S, D = 12, 64, 3
        g = torch.Generator().manual_seed(0)
        full = torch.randn(S, total, D, generator=g).cumsum(0) * 0.1 + torch.randn(S, total, D, generator=g)
        count, off = parallel.shard_chains(total, world, rank)
        local = full[:, off:off + count].contiguous()
        gathered = diagnostics.gather_chains(local)
        ok_gather = torch.equal(gathered, full)
        summ = diagnostics.summarize(local, columns=[0, 2])
        ref = diagnostics.summarize(full, columns=[0, 2], gather=False) if rank == 0 else None
        q.put((rank, ok_gather, summ["min_ess"], summ["max_rhat"], summ["n_chains"],
               None if ref is None else (ref["min_ess"], ref["max_rhat"])))
    finally:
        dist.destroy_process_group()


def test_gloo_two_rank_diagnostics_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort()
    assert all(r[1] for r in results)  # all-gather reproduces the unsharded trace in rank order
    assert results[0][4] == 12 and results[1][4] == 12
    # every rank computes the same pooled diagnostics, equal to the single-process values
    assert results[0][2] == pytest.approx(results[1][2]) and results[0][3] == pytest.approx(results[1][3])
    ref_ess, ref_rhat = results[0][5]
    assert results[0][2] == pytest.approx(ref_ess) and results[0][3] == pytest.approx(ref_rhat)


def test_diagnostics_on_known_series():
    g = torch.Generator().manual_seed(1)
    iid = torch.randn(400, 16, generator=g)
    ess = diagnostics.ess_bulk(iid)
    assert 0.7 * iid.numel() < ess < 1.4 * iid.numel()
    assert diagnostics.split_rhat(iid) < 1.01
    # AR(1) with rho = 0.9: ESS ~ N (1 - rho) / (1 + rho)
    rho, x = 0.9, torch.zeros(2000, 8)
    e = torch.randn(2000, 8, generator=g)
    for t in range(1, 2000):
        x[t] = rho * x[t - 1] + e[t] * (1 - rho ** 2) ** 0.5
    ess = diagnostics.ess_bulk(x)
    expect = x.numel() * (1 - rho) / (1 + rho)
    assert 0.6 * expect < ess < 1.6 * expect
    shifted = iid.clone()
    shifted[:, :8] += 2.0
    assert diagnostics.split_rhat(shifted) > 1.2
