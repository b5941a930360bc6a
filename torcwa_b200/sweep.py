"""Sharding of a batched sweep over the GPUs of one box.

Design points are independent, so the path shards embarrassingly: each rank (one process per GPU,
launched with torch.distributed.run) solves a contiguous slice of the flattened batch with its own
`torcwa_b200.rcwa` objects; the ONLY collective is the final gather of the requested S-parameters
(K complex numbers per point; 32 KB for the 4096-point config -- latency bound, nothing to fuse).
NCCL has no complex dtype, so values travel as real pairs.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_points, rank, world):
    """Contiguous, balanced split: the first (n_points % world) ranks get one extra point."""
    base, extra = divmod(n_points, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_sparams(local, n_points, group=None):
    """local: complex [n_local, K] on this rank's device -> complex [n_points, K] on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    K = local.shape[1]
    n_max = -(-n_points // world)
    rdt = torch.float32 if local.dtype == torch.complex64 else torch.float64
    buf = torch.zeros((n_max, K, 2), dtype=rdt, device=local.device)
    buf[: local.shape[0]] = torch.view_as_real(local.contiguous())
    out = torch.empty((world, n_max, K, 2), dtype=rdt, device=local.device)
    dist.all_gather_into_tensor(out.view(world * n_max, K, 2), buf, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_points, r, world)
        parts.append(torch.view_as_complex(out[r, : hi - lo].contiguous()))
    return torch.cat(parts, dim=0)


def solve_sweep(solve_slice, n_points, group=None):
    """solve_slice(lo, hi) -> complex [hi-lo, K] for this rank's points; returns the gathered
    [n_points, K] tensor on every rank."""
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    lo, hi = shard_bounds(n_points, rank, world)
    return gather_sparams(solve_slice(lo, hi), n_points, group)
