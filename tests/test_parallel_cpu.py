"""CPU (gloo, world_size 2): the host-side logic of the multi-GPU path -- identical cost-balanced
partitions on every rank, full coverage without overlap, and the unique-id rendezvous."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kalign_b200 import _lib, parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _lib.load()
    rng = np.random.default_rng(5)            # same seed everywhere: same task list on every rank
    la = rng.integers(50, 2000, size=777)
    lb = rng.integers(50, 2000, size=777)
    cost = la.astype(np.float64) * lb
    b = parallel.partition(lib, cost, world)
    # every rank must agree on the partition
    t = torch.from_numpy(b.astype(np.int64))
    ref = t.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(t, ref))
    # shards cover the list exactly once
    mine = np.zeros(len(cost), dtype=np.int64)
    mine[b[rank]:b[rank + 1]] = 1
    tot = torch.from_numpy(mine)
    dist.all_reduce(tot)
    covered = bool((tot == 1).all())
    # load balance: my share of the cost
    share = float(cost[b[rank]:b[rank + 1]].sum() / cost.sum())
    # rendezvous payload: rank 0's bytes reach everybody (no GPU: the id itself is a dummy here)
    payload = np.arange(parallel.ID_BYTES, dtype=np.uint8) if rank == 0 else np.zeros(parallel.ID_BYTES, dtype=np.uint8)
    tt = torch.from_numpy(payload.copy())
    dist.broadcast(tt, src=0)
    got = bool((tt.numpy() == np.arange(parallel.ID_BYTES, dtype=np.uint8)).all())
    q.put((rank, same, covered, share, got))
    dist.destroy_process_group()


def test_partition_properties_single_process():
    lib = _lib.load()
    rng = np.random.default_rng(1)
    for n in (0, 1, 2, 7, 100, 5000):
        cost = rng.random(n) * 100 + 1
        for world in (1, 2, 3, 8):
            b = parallel.partition(lib, cost, world)
            assert b[0] == 0 and b[-1] == n
            assert all(b[i] <= b[i + 1] for i in range(world))
            if n >= 50 * world:
                shares = [cost[b[r]:b[r + 1]].sum() / cost.sum() for r in range(world)]
                assert max(shares) < 1.5 / world


def test_two_rank_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, covered, share, got in res:
        assert same and covered and got
        assert 0.35 < share < 0.65


def _ensemble_worker(rank, world, port, q):
    """the host side of an ensemble sharded over `world` processes: every rank resolves the parameters of ITS runs
    (kb200_ensemble_run_params, host code) and the ranks gather them with gloo, as a caller would gather the rows"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_runs, seed = 9, 42
    mine = torch.full((n_runs, 6), -1.0, dtype=torch.float64)
    for k in range(rank, n_runs, world):
        g, e, t, ts, nz = _lib.ensemble_run_params(5.5, 2.0, 1.0, k, seed)
        mine[k] = torch.tensor([k, g, e, t, float(ts), nz], dtype=torch.float64)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    q.put((rank, torch.stack(parts).numpy()))
    dist.destroy_process_group()


def test_ensemble_runs_two_rank_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ensemble_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0], res[1])                       # every rank ends with the same table
    parts = res[0]
    for k in range(9):
        owner = k % world
        assert parts[owner][k][0] == k and parts[1 - owner][k][0] == -1     # each run resolved by exactly one rank
        g, e, t, ts, nz = _lib.ensemble_run_params(5.5, 2.0, 1.0, k, 42)
        assert np.array_equal(parts[owner][k], np.array([k, g, e, t, float(ts), nz]))
