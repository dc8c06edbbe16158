"""Multi-GPU plumbing: the parameter axis is the only parallel axis.

Every output element depends on exactly one parameter's slab (reference loops at
src/ess_rhat.jl:380,517, src/rhat_nested.jl:145), so rank r of W owns the contiguous parameter
range `shard_range(P, r, W)` — one contiguous byte range of the column-major array — and the
only exchange step is the gather of per-parameter scalars to rank 0.  One process per GPU,
`torch.distributed` (NCCL on GPUs; gloo in the CPU tests).  Results are independent of W
(per-parameter computation is batch-invariant).
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) of `total` parameters for `rank` of `world`."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    return rank * total // world, (rank + 1) * total // world


def shard_sizes(total: int, world: int) -> list[int]:
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def gather_params(local, total: int, dst: int = 0, group=None):
    """Gather per-parameter results (a tensor whose LAST dim is this rank's parameter shard) to
    rank `dst`; returns the concatenated tensor there and None elsewhere.  Uneven shards are
    padded to a common width for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(total, world)
    if local.shape[-1] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[-1]} parameters, expected {sizes[rank]}")
    # collectives want equal sizes: pad every shard to the largest one, trim after the gather
    width = max(sizes)
    if local.shape[-1] != width:
        pad = torch.zeros(local.shape[:-1] + (width - local.shape[-1],), dtype=local.dtype, device=local.device)
        local = torch.cat((local, pad), dim=-1)
    local = local.contiguous()
    bufs = None
    if rank == dst:
        bufs = [torch.empty_like(local) for _ in sizes]
    dist.gather(local, bufs, dst=dst, group=group)
    return torch.cat([b[..., :s] for b, s in zip(bufs, sizes)], dim=-1) if rank == dst else None


def sharded_call(fn, samples, group=None, dst: int = 0):
    """Run `fn(shard)` on this rank's parameter shard of `samples` (draws, chains, params) and gather
    the per-parameter results to `dst`.  `fn` may return one array, a tuple of arrays (e.g.
    `(ess, rhat)`: gathered as a stacked `(k, params)` tensor) or a dict of arrays (e.g. `summary`:
    gathered as a dict of `(params,)` tensors).  Returns None on the other ranks."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    total = samples.shape[2]
    lo, hi = shard_range(total, rank, world)
    res = fn(samples[:, :, lo:hi])
    keys = None
    if isinstance(res, dict):
        keys = list(res)
        parts = [res[k] for k in keys]
    elif isinstance(res, (tuple, list)):
        parts = list(res)
    else:
        parts = [res]
    out = torch.stack([torch.as_tensor(p).reshape(-1) for p in parts])
    full = gather_params(out, total, dst=dst, group=group)
    if full is None:
        return None
    if keys is not None:
        return {k: full[i] for i, k in enumerate(keys)}
    if not isinstance(res, (tuple, list)):
        return full[0]
    return full
