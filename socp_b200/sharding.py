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


def gather_results(x, info, nfev, group=None, total=None):
    """All-gather the shards; returns (x[T, P], info[T], nfev[T]) on every rank, T = the global problem count.

    Shards may be ragged (shard_bounds gives the last ranks short or empty blocks when `total` is not a
    multiple of the world size): every rank pads its block to ceil(total/world) rows for the collective and
    the padding is trimmed afterwards.  total=None means equally sized shards (T = world * rows)."""
    pack = pack_results(x, info, nfev)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rows, width = pack.shape
    if world == 1:
        out = pack if total is None else pack[:total]
    else:
        per = rows if total is None else -(-int(total) // world)
        if rows > per:
            raise ValueError("shard of %d rows exceeds ceil(total/world) = %d" % (rows, per))
        if rows < per:
            pad = torch.zeros((per - rows, width), dtype=pack.dtype, device=pack.device)
            pack = torch.cat([pack, pad], dim=0).contiguous()
        out = torch.empty((world * per, width), dtype=pack.dtype, device=pack.device)
        dist.all_gather_into_tensor(out, pack, group=group)
        if total is not None:
            out = out[:int(total)]          # contiguous blocks in rank order: the padding sits at the end
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
