"""Two-GPU test of the one exchange step of the path: the pooled step size of a run whose chains are
sharded over the GPUs of the box (one (sum, count) all-reduce per burn-in transition).
Needs 2 GPUs (`gpurun --gpus 2`); skipped elsewhere."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from pymc4_b200 import _cabi, parallel
        from pymc4_b200._cstructs import PM4_EPS_POOLED, make_adapt, make_nuts_cfg
        from pymc4_b200.ir import lower_model
        from tests import models as M

        ir, _, _ = lower_model(M.radon_model(real=False))
        total, B, S = 96, 30, 10
        count, off = parallel.shard_chains(total, world, rank)
        dm = _cabi.DeviceModel(ir, device=rank)
        comm = _cabi.Communicator(rank, rank, world)
        dm.attach_comm(comm)
        adapt = make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=B)
        out = None
        for rep in range(2):  # twice: the mailbox epochs must keep successive runs apart
            cfg = make_nuts_cfg(count, S, B, step_size=0.1, seed=17, adapt=adapt, chain_offset=off)
            out = dm.nuts(cfg, np.zeros((count, ir.D), np.float32), dtype=np.float32)
            torch.cuda.synchronize()
        eps = out["step_size"].cpu().numpy()
        ref = None
        if rank == 0:  # the unsharded run of all chains on one GPU, no communicator
            dm1 = _cabi.DeviceModel(ir, device=rank)
            cfg = make_nuts_cfg(total, S, B, step_size=0.1, seed=17, adapt=adapt, chain_offset=0)
            o1 = dm1.nuts(cfg, np.zeros((total, ir.D), np.float32), dtype=np.float32)
            ref = (o1["step_size"].cpu().numpy(), o1["leapfrogs"].cpu().numpy())
        q.put((rank, off, count, eps, out["leapfrogs"].cpu().numpy(), ref))
        dist.barrier()
        comm.close()
    finally:
        dist.destroy_process_group()


def test_pooled_step_size_is_shared_across_gpus():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=600) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
    (_, off0, n0, eps0, leaps0, ref), (_, off1, n1, eps1, leaps1, _) = results
    # one epsilon for the whole box, identical on both ranks
    assert np.all(eps0 == eps0[0]) and np.all(eps1 == eps0[0])
    ref_eps, ref_leaps = ref
    # ... and the one the unsharded run adapts (the mean acceptance is the same number up to the order of a
    # double-precision sum; chaotic dynamics may amplify a last-bit difference late in the run)
    np.testing.assert_allclose(eps0[0], ref_eps[0], rtol=2e-3)
    both = np.concatenate([leaps0, leaps1], axis=1)
    assert both.shape == ref_leaps.shape
    assert np.mean(both == ref_leaps) > 0.9


def _worker_dense(rank, world, port, q):
    """The lock-step path (dense GLM likelihood) under a communicator: HMC and NUTS, one epsilon for the box."""
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from pymc4_b200 import _cabi, parallel
        from pymc4_b200._cstructs import PM4_EPS_POOLED, make_adapt, make_hmc_cfg, make_nuts_cfg
        from pymc4_b200.ir import lower_model
        from tests import models as M

        ir, _, _ = lower_model(M.logistic_glm_model(*M.logistic_glm_data(n=1500, p=12)))
        total, B, S = 64, 25, 8
        count, off = parallel.shard_chains(total, world, rank)
        dm = _cabi.DeviceModel(ir, device=rank)
        comm = _cabi.Communicator(rank, rank, world)
        dm.attach_comm(comm)
        adapt = make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=B)
        th0 = np.zeros((count, ir.D), np.float32)
        hmc = dm.hmc(make_hmc_cfg(count, S, B, step_size=0.02, num_leapfrog_steps=4, seed=3, adapt=adapt, chain_offset=off),
                     th0, dtype=np.float32)
        nuts = dm.nuts(make_nuts_cfg(count, S, B, step_size=0.02, seed=4, adapt=adapt, chain_offset=off, max_tree_depth=5),
                       th0, dtype=np.float32)
        torch.cuda.synchronize()
        ref = None
        if rank == 0:
            dm1 = _cabi.DeviceModel(ir, device=rank)
            t0 = np.zeros((total, ir.D), np.float32)
            h1 = dm1.hmc(make_hmc_cfg(total, S, B, step_size=0.02, num_leapfrog_steps=4, seed=3, adapt=adapt), t0, dtype=np.float32)
            n1 = dm1.nuts(make_nuts_cfg(total, S, B, step_size=0.02, seed=4, adapt=adapt, max_tree_depth=5), t0, dtype=np.float32)
            ref = (h1["step_size"].cpu().numpy(), n1["step_size"].cpu().numpy())
        q.put((rank, hmc["step_size"].cpu().numpy(), nuts["step_size"].cpu().numpy(), ref))
        dist.barrier()
        comm.close()
    finally:
        dist.destroy_process_group()


def test_pooled_step_size_is_shared_across_gpus_on_the_lockstep_path():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_dense, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=600) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
    (_, h0, n0, ref), (_, h1, n1, _) = results
    assert np.all(h0 == h0[0]) and np.all(h1 == h0[0]) and np.all(n0 == n0[0]) and np.all(n1 == n0[0])
    np.testing.assert_allclose(h0[0], ref[0][0], rtol=0.05)
    np.testing.assert_allclose(n0[0], ref[1][0], rtol=0.05)
