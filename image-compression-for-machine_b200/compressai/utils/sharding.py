"""Batch sharding across ranks (one process per GPU).  Images are independent, so the data path needs no
collective: each rank codes a contiguous slice of the batch; only the variable-length byte strings and the
timing are gathered on the host.  Works with any torch.distributed backend (NCCL on the GPU
box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced slice [lo, hi) of n_items for `rank` (the first n_items % world ranks get one more)."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, device="cpu"):
    """Max of a python float over all ranks (the multi-GPU timing rule: report the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_strings(strings, dst=0):
    """Gather per-image byte strings (`[[y_0..], [z_0..]]` of this rank's slice) on rank `dst`, in batch order.

    Returns the concatenated [[y...], [z...]] on `dst`, None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return strings
    world, rank = dist.get_world_size(), dist.get_rank()
    bucket = [None] * world if rank == dst else None
    dist.gather_object(strings, bucket, dst=dst)
    if rank != dst:
        return None
    return [[s for part in bucket for s in part[0]], [s for part in bucket for s in part[1]]]


def scatter_strings(strings, n_items, src=0):
    """Inverse of gather_strings: every rank receives the strings of its own slice."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return strings
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = None
    if rank == src:
        parts = []
        for r in range(world):
            lo, hi = shard_range(n_items, r, world)
            parts.append([strings[0][lo:hi], strings[1][lo:hi]])
    out = [None]
    dist.scatter_object_list(out, parts, src=src)
    return out[0]
