"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the rendezvous and the
timing reductions, the library's own NCCL communicator for the data path (anchor position maps,
per-level coded paths and merged profiles).  The data path never goes through torch."""
import ctypes as C

import numpy as np

ID_BYTES = 128


def exchange_unique_id(lib, rank, dist, device=None):
    """rank 0 creates the NCCL unique id (kb200_comm_unique_id), everybody receives it through
    torch.distributed (works with the nccl and the gloo backend)."""
    import torch
    buf = np.zeros(ID_BYTES, dtype=np.uint8)
    if rank == 0 and lib is not None:
        if lib.kb200_comm_unique_id(buf.ctypes.data_as(C.c_void_p), ID_BYTES) != 0:
            raise RuntimeError("kb200_comm_unique_id failed")
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    return t.cpu().numpy().astype(np.uint8).copy()


def attach(ctx, rank, world, id_bytes):
    """attach the context to the communicator (kb200_ctx_comm_init)"""
    id_bytes = np.ascontiguousarray(id_bytes, dtype=np.uint8)
    if ctx.lib.kb200_ctx_comm_init(ctx.h, rank, world, id_bytes.ctypes.data_as(C.c_void_p), len(id_bytes)) != 0:
        raise RuntimeError("kb200_ctx_comm_init failed")


def partition(lib, cost, world):
    """contiguous cost-balanced shards (kb200_partition): bounds[0..world]"""
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    b = np.zeros(world + 1, dtype=np.int32)
    if lib.kb200_partition(cost, len(cost), world, b) != 0:
        raise RuntimeError("kb200_partition failed")
    return b
