"""Multi-GPU plumbing: chains shard embarrassingly across ranks (one process per GPU); no collective runs while
sampling.  The only exchange is the optional assembly of ``draws_out`` on every rank — one all-gather of each rank's
chain-major block (NCCL over NVLink on GPUs; gloo in the CPU tests).  Because Philox counters and MT19937 seeds use
GLOBAL chain ids (``chain_offset``), results are independent of the number of ranks."""
import torch
import torch.distributed as dist


def chain_shard(n_chains_total, rank, world_size):
    """Contiguous partition: rank g owns chains [first, first+count).  Remainders go to the lowest ranks."""
    base, rem = divmod(int(n_chains_total), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def all_gather_draws(local_draws, n_chains_total, group=None):
    """local_draws: [count_r, n_keep, d] tensor of this rank's chains -> [n_chains_total, n_keep, d] on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [chain_shard(n_chains_total, r, world)[1] for r in range(world)]
    assert local_draws.shape[0] == counts[rank]
    tail = tuple(local_draws.shape[1:])
    if len(set(counts)) == 1:
        out = torch.empty((n_chains_total,) + tail, dtype=local_draws.dtype, device=local_draws.device)
        dist.all_gather_into_tensor(out, local_draws.contiguous(), group=group)
        return out
    cmax = max(counts)
    padded = torch.zeros((cmax,) + tail, dtype=local_draws.dtype, device=local_draws.device)
    padded[: counts[rank]] = local_draws
    buf = torch.empty((world * cmax,) + tail, dtype=local_draws.dtype, device=local_draws.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    parts = [buf[r * cmax: r * cmax + counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0)


def all_reduce_max(value, device, group=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
