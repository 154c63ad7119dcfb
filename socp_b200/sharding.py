"""Multi-GPU layout of a batch: independent problems, contiguous block per rank, no data-path
collective.  The only exchange is the final gather of the converged unknowns and
solver statistics, over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Contiguous block of ceil(total/world) problems for `rank` (the last ranks may be short or empty)."""
    per = -(-total // world)
    lo = min(rank * per, total)
    return lo, min(lo + per, total)


def pack_results(x, info, nfev):
    """One float64 row per problem: [x[0..P), info, nfev]."""
    return torch.cat([x, info.to(torch.float64)[:, None], nfev.to(torch.float64)[:, None]], dim=1).contiguous()


def gather_results(x, info, nfev, group=None):
    """All-gather equally sized shards; returns (x[W*B, P], info[W*B], nfev[W*B]) on every rank."""
    pack = pack_results(x, info, nfev)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        out = pack
    else:
        out = torch.empty((world * pack.shape[0], pack.shape[1]), dtype=pack.dtype, device=pack.device)
        dist.all_gather_into_tensor(out, pack, group=group)
    P = x.shape[1]
    return out[:, :P], out[:, P].to(torch.int32), out[:, P + 1].to(torch.int32)


def reduce_sum(values, device, group=None):
    """Sum a list of python floats over the ranks (converged counts, RK4 steps, ...)."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(v) for v in t.tolist()]


def reduce_max(value, device, group=None):
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
