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


# ------------------------------------------------------------------------------------------ chunking within one GPU
def points_per_call(order, free_bytes=None, device=None, safety=0.8):
    """How many design points one `rcwa` object should carry on this GPU.  The peak of a patterned layer is about 11
    live n x n complex128 matrices per point (layer S-matrix: W, Q, S11, S21 + 6 workspace; star product: 2 + 4
    outputs + 5 workspace), n = 2 (2 ox + 1)(2 oy + 1): 0.65 GB per point at order 15, i.e. 128
    points on a 180 GB B200.  Throughput is flat from 128 points on (measured), so the result is capped there."""
    ox, oy = (order, order) if isinstance(order, int) else (int(order[0]), int(order[1]))
    n = 2 * (2 * ox + 1) * (2 * oy + 1)
    if free_bytes is None:
        free_bytes, _ = torch.cuda.mem_get_info(device)
    per_point = 11 * n * n * 16 * 1.15
    return max(1, min(128, int(safety * free_bytes / per_point)))


def chunked(solve_slice, chunk):
    """Wrap solve_slice(lo, hi) -> [hi-lo, K] so that it is called on pieces of at most `chunk` points (one `rcwa`
    object each) and the pieces are concatenated: the way to run a 512-point sweep on one GPU."""
    def run(lo, hi):
        parts = [solve_slice(a, min(a + chunk, hi)) for a in range(lo, hi, chunk)]
        return torch.cat(parts, dim=0) if parts else solve_slice(lo, hi)
    return run
