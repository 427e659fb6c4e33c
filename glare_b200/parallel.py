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


class AsyncGather:
    """all_gather of each rank's result batch on a SIDE stream: the NVLink transfer of batch i overlaps the kernels of batch i + 1 (the
    forward pass itself has no collective).  ``submit`` snapshots the local result into one of two private slots on the compute stream (the
    caller's tensor -- e.g. a CUDA graph's static output -- may be overwritten right away), the collective runs on the side stream;
    ``wait`` makes the compute stream wait for everything submitted so far.  Without an initialised process group it is a pass-through."""

    def __init__(self, group=None):
        self.group = group
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.on else 1
        self.side = torch.cuda.Stream() if self.on and torch.cuda.is_available() else None     # CPU tensors (gloo tests): synchronous
        self.slots, self.outs, self.done, self.i = [None, None], [None, None], [None, None], 0

    def submit(self, local):
        if not self.on:
            return local
        i, self.i = self.i, self.i ^ 1
        if self.side is None:
            out = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
            return out
        cur = torch.cuda.current_stream()
        if self.slots[i] is None or self.slots[i].shape != local.shape or self.slots[i].dtype != local.dtype:
            self.slots[i] = torch.empty_like(local)
            self.outs[i] = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        if self.done[i] is not None:
            cur.wait_event(self.done[i])                  # the collective that last read this slot has finished
        self.slots[i].copy_(local)
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            dist.all_gather_into_tensor(self.outs[i], self.slots[i], group=self.group)
            self.done[i] = torch.cuda.Event()
            self.done[i].record(self.side)
        return self.outs[i]

    def wait(self):
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)


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
