"""Data-parallel plumbing.  Frame-pairs are independent units: the global batch is split
contiguously over ranks (both legs of a pair stay together -- the correlation needs them), RoI image indices
are shard-local, and the eval forward has NO data-path collective.  The only exchanges are the timing
reduction (MAX over ranks) and, if wanted, a gather of results onto one rank.  The reference's equivalent is
nn.DataParallel's scatter on dim 0 (trainval_net.py:311,365).  One process per GPU, torch.distributed (NCCL on
GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_pairs, rank, world):
    """pairs [lo, hi) owned by `rank`: contiguous, balanced to +-1"""
    return n_pairs * rank // world, n_pairs * (rank + 1) // world


def localize_rois(rois, first_pair):
    """global image index in column 0 -> shard-local"""
    out = rois.clone()
    out[:, 0] -= first_pair
    return out


def max_over_ranks(t):
    """device/host timings: the job is as slow as its slowest rank"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def throughput(units_per_rank, ms, world):
    """whole-job units per second"""
    return world * units_per_rank / (ms / 1e3)


def gather_pairs(part, dst=0):
    """concatenate per-shard results (dim 0 = pairs, in rank order) on rank `dst`; None elsewhere"""
    world = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.int64, device=part.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([part.size(0)], dtype=torch.int64, device=part.device))
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros((mx,) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device)
    pad[: part.size(0)] = part
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if dist.get_rank() != dst:
        return None
    return torch.cat([b[: int(s)] for b, s in zip(bufs, sizes)], 0)


def allreduce_gradients(params, bucket_bytes=64 << 20):
    """The ONE collective of the training step: mean of the gradients of the trainable
    parameters over ranks, in a few flat buckets (reverse parameter order = the order backward produces
    them).  Replaces nn.DataParallel's reduce-onto-GPU-0 (trainval_net.py:311); with equal shard sizes the
    result equals the gradient of the mean loss over the global batch (trainval_net.py:367-368)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    grads = [p.grad for p in reversed(list(params)) if p.requires_grad and p.grad is not None]
    n_buckets, i = 0, 0
    while i < len(grads):
        j, size = i, 0
        while j < len(grads) and (size == 0 or size + grads[j].numel() * grads[j].element_size() <= bucket_bytes):
            size += grads[j].numel() * grads[j].element_size()
            j += 1
        flat = torch.cat([g.reshape(-1) for g in grads[i:j]])
        dist.all_reduce(flat)
        flat.div_(world)
        off = 0
        for g in grads[i:j]:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_buckets += 1
        i = j
    return n_buckets
