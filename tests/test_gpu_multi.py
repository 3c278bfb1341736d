"""2-GPU parity (run with `gpurun --gpus 2`): the level-sharded multi-GPU path must return exactly
the single-GPU alignment (which is the reference's) on every rank."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from kalign_b200 import _lib, parallel, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = _lib.Context(rank)
    uid = parallel.exchange_unique_id(ctx.lib, rank, dist, device=torch.device("cuda", rank))
    parallel.attach(ctx, rank, world, uid)
    out = []
    for seqs, type_, K in ((synth.family(150, 200, synth.PROTEIN, seed=41), 8, 5),
                           (synth.family(90, 300, synth.RNA, seed=42), 2, 5),
                           (synth.family(64, 120, synth.PROTEIN, seed=43), 8, 0)):
        out.append(ctx.kalign(seqs, n_threads=2, type_=type_, consistency=K, weight=2.0))
    st = ctx.stats()
    q.put((rank, out, st["n_collectives"]))
    dist.barrier()
    ctx.lib.kb200_ctx_comm_destroy(ctx.h)
    ctx.close()
    dist.destroy_process_group()


def test_two_gpus_identical_to_one():
    from kalign_b200 import _lib, synth
    if _lib.load().kb200_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx1 = _lib.Context(0)
    want = []
    for seqs, type_, K in ((synth.family(150, 200, synth.PROTEIN, seed=41), 8, 5),
                           (synth.family(90, 300, synth.RNA, seed=42), 2, 5),
                           (synth.family(64, 120, synth.PROTEIN, seed=43), 8, 0)):
        want.append(ctx1.kalign(seqs, n_threads=2, type_=type_, consistency=K, weight=2.0))
    ctx1.close()
    world = 2
    port = _free_port()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, out, ncoll in res:
        assert ncoll > 0
        assert out == want, "rank %d differs" % rank


def _ensemble_worker(rank, world, q):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import gen_golden_seeded as G
    from kalign_b200 import _lib
    ctx = _lib.Context(rank)                       # one process per GPU; no communicator: the runs share nothing
    out = {}
    for fk, (n_runs, seed) in sorted(G.ENSEMBLE.items()):
        seqs, type_ = G.families()[fk]
        out[fk] = ctx.ensemble_runs(seqs, n_runs, seed=seed, rank=rank, world=world, n_threads=2, type_=type_)
    q.put((rank, out))
    ctx.close()


def test_ensemble_runs_sharded_over_two_gpus():
    """the ensemble's independent runs batched across GPUs -- rank r computes the runs k % 2 == r on
    its own GPU, nothing is exchanged on the data path; together they are the reference's runs (goldens of
    kalign_run_seeded with resolve_run_params' parameters, tests/golden/seeded.npz)"""
    import sys
    from kalign_b200 import _lib
    if _lib.load().kb200_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import gen_golden_seeded as G
    Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "seeded.npz"))
    world = 2
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_ensemble_worker, args=(r, world, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for fk, (n_runs, seed) in G.ENSEMBLE.items():
        for k in range(n_runs):
            owner = k % world
            assert k in res[owner][fk] and k not in res[1 - owner][fk]
            assert res[owner][fk][k] == [str(x) for x in Z["ens_%s_%d" % (fk, k)]], (fk, k)
