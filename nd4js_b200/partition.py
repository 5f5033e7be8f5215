"""Batch partitioner: the independent leading batch index is split into contiguous, near-equal
ranges (row-major => each range is one contiguous slice of every operand and result).

Inside one process libnd4b shards over its context's devices the same way (nd4b_api.cu,
run_pipeline); across processes (one per GPU under torchrun) these helpers give each rank its
range.  No collective is on the data path; `gather_shards` exists for callers that want the
whole result on every rank (NCCL on GPU tensors, gloo on CPU tensors)."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """[begin, end) of the flattened batch index owned by `rank`."""
    return total * rank // world, total * (rank + 1) // world


def gather_shards(local, total):
    """All-gathers contiguous shards (dim 0) produced with shard_range into the full batch."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    if min(sizes) == max(sizes):  # equal shards: one collective straight into the result, no staging copies
        out = local.new_empty((total,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def max_over_ranks(value, device=None):
    """Timing reduction of the bench contract: the job is as slow as its slowest rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
