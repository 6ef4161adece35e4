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


def all_gather_summary(chain_mean, chain_var, n_keep, n_chains_total, group=None):
    """Posterior summaries over ALL ranks' chains without gathering any draws ("summary reductions fused
    into the all-gather epilogue"): every rank passes the per-chain means / sample variances of its shard ([count_r, d] each,
    e.g. from ``api.summarize(..., per_chain=True)``, which reduces the device-resident draws_out where it lies); the
    O(n_chains * d) statistics are all-gathered and combined the same way the single-GPU kernel does.  Returns
    (mean[d], var[d], rhat[d]) of the n_chains_total * n_keep draws on every rank."""
    cm = all_gather_draws(chain_mean.contiguous(), n_chains_total, group=group).to(torch.float64)
    cv = all_gather_draws(chain_var.contiguous(), n_chains_total, group=group).to(torch.float64)
    C, n = float(n_chains_total), float(n_keep)
    shift = cm[0]
    e = cm - shift
    t1, t2, tw = e.sum(dim=0), (e * e).sum(dim=0), cv.sum(dim=0)
    mean = shift + t1 / C
    sb = torch.clamp(t2 - t1 * (t1 / C), min=0.0)                 # sum_c (m_c - m)^2
    W = tw / C
    Bn = sb / (C - 1.0) if n_chains_total > 1 else torch.full_like(W, float("nan"))
    var = ((n - 1.0) * tw + n * sb) / (C * n - 1.0)
    rhat = torch.sqrt(((n - 1.0) / n * W + Bn) / W)
    return mean, var, rhat
