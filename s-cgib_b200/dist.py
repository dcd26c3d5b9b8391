"""Data parallelism over molecules: one process per GPU, each rank trains on its own
mini-batch shard (BatchNorm statistics, recon and contrastive losses stay per rank, exactly as if the reference ran at
the per-rank batch size); the only exchange is ONE all-reduce of the flat gradient buffer per step (NCCL over
NVLink 5 / NVSwitch), followed by the fused Adam with grad_scale = 1 / world_size."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).  Returns
    (rank, local_rank, world)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, local_rank, world


def shard_range(num_graphs: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of a global batch: rank r takes graphs r*B_local .. (r+1)*B_local (the remainder is
    spread over the first ranks)."""
    base, rem = divmod(num_graphs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_grads_(flat_grads: torch.Tensor, world: int) -> float:
    """Sum-all-reduce the flat gradient buffer in place; returns the scale the optimiser applies (1/world)."""
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return 1.0 / world


def broadcast_params_(flat_params: torch.Tensor, src: int = 0):
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat_params, src)
