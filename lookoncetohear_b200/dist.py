"""Multi-GPU plumbing.  The path shards embarrassingly: mixtures/streams are independent, so the
only communication is a one-off broadcast of weights and embeddings from rank 0 (NCCL over
NVLink on GPUs; gloo on CPU for the tests) and, optionally, a gather of per-rank results."""
import os

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous slice [lo, hi) of n_items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / MASTER_*); returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"),
                                rank=rank, world_size=world)
    return rank, world


def broadcast_module(module, src=0):
    """Broadcast every parameter and buffer of `module` from rank `src`, then mark the engine's
    packed weights dirty so the next call re-uploads them."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src)
    if hasattr(module, "refresh_weights"):
        module.refresh_weights()
    return module


def broadcast_tensor(t, src=0):
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src=src)
    return t


def gather_counts(value, device="cpu"):
    """All-gather one float per rank (e.g. frames processed, elapsed ms) -> list."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return [float(value)]
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(o[0]) for o in out]
