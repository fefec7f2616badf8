"""CPU, world_size 2 over gloo: the multi-GPU path of bench.py is pure batch sharding -- every rank owns a contiguous
slice of the environments, there is no data-path collective, and the reported time is the max over ranks.  The shard
arithmetic and the reduction are exercised here with the oracle standing in for the kernel (checker role only)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def shard(batch, world, rank):
    per = batch // world
    return rank * per, per


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O

    B, T = 64, 50
    rng = np.random.default_rng(1234)  # same global problem on every rank
    state = [rng.uniform(-1, 1, B), rng.uniform(-1, 1, B)]
    acts = rng.uniform(-1, 1, (B, T, 1))
    lo, n = shard(B, world, rank)
    orc = O.Oracle("mass_spring_damper", solver=O.TSIT5)
    mine = orc.sim_ahead([s[lo:lo + n] for s in state], acts[lo:lo + n], 1e-4, 1e-4)["obs"]
    # no collective on the data path; results are only gathered for checking
    parts = [torch.zeros_like(torch.as_tensor(mine)) for _ in range(world)]
    dist.all_gather(parts, torch.as_tensor(mine))
    t = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)  # stand-in for per-rank CUDA-event time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = orc.sim_ahead(state, acts, 1e-4, 1e-4)["obs"]
        q.put((bool(np.array_equal(torch.cat(parts).numpy(), full)), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_batch_sharding_world_size_2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same, "sharded trajectories must equal the unsharded ones bit for bit (environments are independent)"
    assert abs(tmax - 0.2) < 1e-12  # max over ranks


def test_shard_arithmetic():
    covered = []
    for r in range(8):
        lo, n = shard(1 << 20, 8, r)
        covered.append((lo, lo + n))
    assert covered[0][0] == 0 and covered[-1][1] == 1 << 20
    assert all(covered[i][1] == covered[i + 1][0] for i in range(7))
