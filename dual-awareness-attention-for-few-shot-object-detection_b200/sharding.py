"""Episode sharding across the GPUs of one box.  One episode (a query + its support set) never reads
another episode's data (every op of dana.py:87-220 is per batch element), so ranks get contiguous
slices of the episode list and the forward path has NO collective.  torch.distributed is used only
to agree on the step time (max over ranks) and the unit count (sum over ranks) for reporting."""
import torch


def shard_range(n_units, rank, world):
    """[begin, end) of the units owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(n_units, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def aggregate_throughput(units, seconds, device="cuda"):
    """Whole-job throughput = (sum of units over ranks) / (max of seconds over ranks).
    Returns (value, max_seconds, total_units).  Works without an initialised process group (world 1)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
        u = torch.tensor([float(units)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        seconds, units = float(t.item()), float(u.item())
    return (units / seconds if seconds > 0 else 0.0), float(seconds), int(round(units))
