"""Sharding a problem batch over the GPUs of one box.

Problems are independent: each rank solves a contiguous shard
with no communication, and ONE collective at the end gathers the per-problem
costs (or, for multi-start batches, the best start of every scene).  Backend is
whatever ``torch.distributed`` was initialised with: NCCL over NVLink on the
GPU box, gloo in the CPU tests.
"""

import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [lo, hi) of ``total`` items owned by ``rank``; shards differ by
    at most one item."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_scenes(num_scenes, per_scene, rank, world):
    """Multi-start batches are sharded by scene so that every argmin group stays
    on one GPU: returns the scene range and the problem range of ``rank``."""
    s_lo, s_hi = shard_range(num_scenes, rank, world)
    return (s_lo, s_hi), (s_lo * per_scene, s_hi * per_scene)


def gather_costs(local_costs, counts=None):
    """all_gather of the per-problem ``traj_costs`` -> one tensor in problem order.
    ``counts`` (problems per rank) is needed only for uneven shards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_costs.clone()
    world = dist.get_world_size()
    if counts is None or len(set(counts)) == 1:
        out = torch.empty(world * local_costs.numel(), dtype=local_costs.dtype, device=local_costs.device)
        dist.all_gather_into_tensor(out, local_costs.contiguous())
        return out
    width = max(counts)
    padded = torch.full((width,), float("inf"), dtype=local_costs.dtype, device=local_costs.device)
    padded[:local_costs.numel()] = local_costs
    out = torch.empty(world * width, dtype=local_costs.dtype, device=local_costs.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * width:r * width + c] for r, c in enumerate(counts)])


def gather_best(local_min, local_arg, problem_offset, counts=None):
    """Final argmin gather of a scene-sharded multi-start batch: every rank
    contributes (min cost, global problem index) of its scenes; returns both for
    all scenes, in scene order, on every rank.  ``counts`` (scenes per rank, e.g.
    from ``shard_range``) is needed when the shards are uneven: the collective
    wants equal contributions, so every rank pads to the widest shard with
    (inf, -1) and the padding is dropped after the gather."""
    arg = torch.where(local_arg >= 0, local_arg.to(torch.int64) + int(problem_offset),
                      torch.full_like(local_arg, -1, dtype=torch.int64))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_min.clone(), arg
    world = dist.get_world_size()
    if counts is not None and len(counts) != world:
        raise ValueError(f"counts has {len(counts)} entries for {world} ranks")
    if counts is not None and counts[dist.get_rank()] != local_min.numel():
        raise ValueError("counts does not match this rank's number of scenes")
    # one collective: pack (cost, index) as two float64 columns (indices < 2^53 are exact)
    packed = torch.stack([local_min.to(torch.float64), arg.to(torch.float64)], dim=1).contiguous()
    width = packed.shape[0] if counts is None or len(set(counts)) == 1 else max(counts)
    if width != packed.shape[0]:
        pad = torch.empty((width - packed.shape[0], 2), dtype=torch.float64, device=packed.device)
        pad[:, 0] = float("inf")
        pad[:, 1] = -1.0
        packed = torch.cat([packed, pad])
    out = torch.empty((world * width, 2), dtype=torch.float64, device=packed.device)
    dist.all_gather_into_tensor(out, packed)
    if width != local_min.numel() or (counts is not None and len(set(counts)) > 1):
        out = torch.cat([out[r * width:r * width + c] for r, c in enumerate(counts)])
    return out[:, 0].contiguous(), out[:, 1].to(torch.int64)


def global_best(costs):
    """(min, argmin) over a gathered cost vector, ignoring non-finite entries."""
    safe = torch.where(torch.isfinite(costs), costs, torch.full_like(costs, float("inf")))
    m, i = safe.min(dim=0)
    return m, i
