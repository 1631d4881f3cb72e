"""Batch sharding across GPUs: one process per GPU, independent chains partitioned in contiguous blocks, and a
single exchange step -- an all-gather of the per-chain scalars (overlaps / norms).  No collective touches the cores.

Works with any torch.distributed backend: "nccl" over NVLink on the GPU box, "gloo" in the CPU tests of the host logic.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun) if needed; returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kwargs)
    return rank, world


def shard_bounds(total, rank, world):
    """Contiguous block [lo, hi) of `total` independent units owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d out of range for world %d" % (rank, world))
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(total, world):
    return [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]


def gather_scalars(local, total):
    """All-gather the per-unit scalars of every rank into one (total,) tensor, ordered like the unsharded batch.
    `local` holds this rank's block (shard_bounds).  One collective; message = 8 bytes per unit."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(total, world)
    assert local.numel() == sizes[rank], (local.numel(), sizes[rank])
    if len(set(sizes)) == 1:
        out = torch.empty(total, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    parts = [torch.empty(pad, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
