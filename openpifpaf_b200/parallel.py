"""Multi-GPU plumbing for the hot path: images are independent units, so the batch is sharded
contiguously across ranks (one process per GPU) and there is NO collective on the data path
(the reference's DataParallel scatter/gather, predictor.py:33-37, is not reproduced).
The only collectives are the barrier / max-reduce that time a multi-rank run and an optional
gather of per-rank annotation counts.  Works with the nccl (GPU) and gloo (CPU tests) backends."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous split: rank r gets [start, stop) with sizes differing by at most one."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(images, rank, world):
    start, stop = shard_range(images.shape[0], rank, world)
    return images[start:stop]


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device='cpu'):
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(local_counts, device='cpu'):
    """All ranks' per-image annotation counts, in global image order (ranks may hold unequal shards)."""
    local = torch.as_tensor(local_counts, dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()):
        return local.cpu().tolist()
    world = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.numel()], dtype=torch.int64, device=device))
    m = int(max(int(s.item()) for s in sizes))
    padded = torch.full((m,), -1, dtype=torch.int64, device=device)
    padded[:local.numel()] = local
    out = [torch.empty(m, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(out, padded)
    res = []
    for s, o in zip(sizes, out):
        res += o[:int(s.item())].cpu().tolist()
    return res
