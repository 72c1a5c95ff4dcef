"""Process-group plumbing for the multi-GPU path (torch.distributed is plumbing only: the data-path
exchanges -- migration, ghost layers, mesh reduction -- are NCCL calls inside libp3m_b200.so)."""
from __future__ import annotations

import os

import numpy as np


def init_process_group(backend=None):
    """One process per GPU, launched by torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* in the env)."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Rank `src` supplies `payload`; every rank returns it (used for the NCCL unique id)."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    if dist.get_rank() == src:
        assert payload is not None and len(payload) == nbytes
        t = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tolist())


def create_context(params, capi=None):
    """Collective: every rank gets a p3m context bound to its GPU and joined in one NCCL communicator."""
    import torch.distributed as dist
    if capi is None:
        from . import capi
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return capi.Context(params)
    uid = capi.comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 128, 0)
    return capi.Context(params, unique_id=uid, rank=rank, nranks=world)


def owner_of(z_code, cuts, layer_size):
    """numpy mirror of layer_owner() (csrc/common.cuh): rank owning code-unit coordinate z."""
    layer = np.floor(np.asarray(z_code, np.float64) / layer_size).astype(np.int64)
    owner = np.zeros(layer.shape, np.int64)
    for k in range(1, len(cuts) - 1):
        owner += layer >= cuts[k]
    return owner
