"""Multi-GPU plumbing: chains are independent, so ranks own contiguous blocks of GLOBAL chain ids
(Philox is keyed by the global id, results do not depend on the GPU count) and exchange nothing
while sampling.  The only collective is the final reduction of pooled moments and counters
(SURVEY.md section 8(e)): two all-reduces of 2d+7 doubles (latency-bound; NCCL over NVLink on the
GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard(n_chains, rank, world):
    """Contiguous block [start, stop) of global chain ids owned by `rank`."""
    base, rem = divmod(int(n_chains), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_pooled(pooled, d, group=None):
    """pooled: tensor [2d+7] = (n, mean[d], M2[d], 6 counters) of this rank (ChainBatch.pooled()).
    Returns the same layout merged over all ranks with Chan's formula:
      pass 1: sum of (n, n*mean, counters)  -> global n, mean
      pass 2: sum of M2_r + n_r*(mean_r - mean)^2."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return pooled.clone()
    n = pooled[0:1]
    mean = pooled[1:1 + d]
    m2 = pooled[1 + d:1 + 2 * d]
    first = torch.cat([n, n * mean, pooled[1 + 2 * d:]])
    dist.all_reduce(first, op=dist.ReduceOp.SUM, group=group)
    n_tot = first[0:1]
    mean_tot = torch.where(n_tot > 0, first[1:1 + d] / torch.clamp(n_tot, min=1.0), torch.zeros_like(mean))
    second = m2 + n * (mean - mean_tot) ** 2
    dist.all_reduce(second, op=dist.ReduceOp.SUM, group=group)
    return torch.cat([n_tot, mean_tot, second, first[1 + d:]])


def max_over_ranks(x, device, group=None):
    """Device-timed milliseconds -> max over ranks (the number a multi-GPU run reports)."""
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
