"""Multi-GPU data path: images are independent units, so the batch is split across ranks (one process per GPU,
torchrun) with NO collective inside the forward pass; one all_gather of the enhanced images at the end
(SURVEY.md 8e).  The reference's nn.DataParallel scatter / replicate / gather per step
(VQLLFLOWD_model.py:72-75) is replaced by this static split.

Works with any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced split of n units: the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def enhance_sharded(fn, images, group=None):
    """Every rank holds the same ``images`` [N,...] (or at least its own slice); rank r runs ``fn`` on its slice
    and all ranks receive the full result in the original order.  ``fn(x) -> y`` with y.shape[0] == x.shape[0]."""
    if not (dist.is_available() and dist.is_initialized()):
        return fn(images)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = images.shape[0]
    s, e = shard_range(n, rank, world)
    local = fn(images[s:e])
    per = (n + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: e - s] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        rs, re = shard_range(n, r, world)
        parts.append(out[r * per: r * per + (re - rs)])
    return torch.cat(parts, 0)


def allreduce_gradients(grads, group=None):
    """Data-parallel stage-2 training: every rank computes the gradients of its share of the batch
    (glare_b200.encoder_train.stage2_step); this averages them over the ranks with ONE collective over a flat fp32 bucket
    (27.7 M parameters = 111 MB, a single NCCL all-reduce over NVLink / NVSwitch) -- what DistributedDataParallel would do for
    the reference's nll.mean().backward() (LLFlow_model.py:215-232).  ``grads``: {state-dict key: tensor}, identical key sets on
    every rank; returns the same dictionary with averaged tensors (views into the bucket)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grads
    keys = sorted(grads)
    flat = torch.cat([grads[k].reshape(-1).float() for k in keys])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    out, off = {}, 0
    for k in keys:
        n = grads[k].numel()
        out[k] = flat[off:off + n].view(grads[k].shape)
        off += n
    return out
