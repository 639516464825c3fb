"""Candidate sharding helpers for multi-GPU CEM (one process per GPU, torch.distributed)."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_total, rank=None, world_size=None):
    """[first, last) global candidate ids owned by `rank`; n_total must divide evenly (weights are
    replicated, candidates are independent until the elite selection)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    if n_total % world_size:
        raise ValueError("number of candidates (%d) must be a multiple of the world size (%d)" % (n_total, world_size))
    per = n_total // world_size
    return rank * per, (rank + 1) * per


def gather_costs(cost_local):
    """The one exchange step of a CEM iteration: all-gather of the per-candidate costs (rank-major order ==
    global candidate id order).  Works on NCCL (cuda tensors) and gloo (cpu tensors)."""
    r, w = world()
    if w == 1:
        return cost_local
    out = torch.empty(cost_local.shape[0] * w, dtype=cost_local.dtype, device=cost_local.device)
    dist.all_gather_into_tensor(out, cost_local.contiguous())
    return out
