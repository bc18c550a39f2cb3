"""Data-parallel host logic of the path: the op shards by batch item (every index the kernels use is
prefixed by the batch index — /root/reference/projects/mmdet3d_plugin/ops/src/
deformable_aggregation_cuda.cu:165,:174), so ranks never exchange activations.  The one collective
of a training step is the gradient all-reduce of the replicated module parameters, which the
reference gets from MMDistributedDataParallel (projects/mmdet3d_plugin/apis/mmdet_train.py); here
it is a flat-bucket all-reduce that can run on a side stream while the next step computes.

One process per GPU; `torch.distributed` (NCCL on the GPU box, gloo in the CPU tests) is plumbing.
"""
import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard_batch", "GradBucket"]


def shard_range(batch_size, world_size, rank):
    """Contiguous [lo, hi) slice of the batch owned by `rank`; the first batch_size % world_size
    ranks get one extra item.  Every item is owned by exactly one rank."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(batch_size, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors, world_size, rank):
    """Slices every tensor of a dict along dim 0 (batch); tensors without a batch dim (the level
    tables spatial_shape / scale_start_index) are passed through untouched."""
    any_batched = [t for k, t in tensors.items() if k not in ("spatial_shape", "scale_start_index")
                   and isinstance(t, torch.Tensor)]
    lo, hi = shard_range(any_batched[0].shape[0], world_size, rank)
    out = {}
    for k, t in tensors.items():
        if isinstance(t, torch.Tensor) and k not in ("spatial_shape", "scale_start_index"):
            out[k] = t[lo:hi].contiguous()
        else:
            out[k] = t
    return out


class GradBucket:
    """All parameter gradients of a module in ONE flat buffer, averaged across ranks with a single
    all-reduce (DFA: 247,495 fp32 = 0.99 MB per layer — a latency-bound message, so one launch
    instead of one per tensor).  With a CUDA `comm_stream` the all-reduce is enqueued behind an event
    on the compute stream and overlaps whatever the compute stream does next; `wait()` joins.

    alias_grads=True makes every parameter's .grad a VIEW of the flat buffer (what DDP does with its
    buckets): autograd then accumulates straight into the bucket and pack() / unpack() copy nothing.
    On NCCL the mean is taken by the collective itself (ReduceOp.AVG: no separate division kernel)."""

    def __init__(self, params, group=None, comm_stream=None, alias_grads=False):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.comm_stream = comm_stream
        self.alias_grads = alias_grads
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, device=ref.device, dtype=ref.dtype)
        self.views, o = [], 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()
        if alias_grads:
            for p, v in zip(self.params, self.views):
                if p.grad is not None:
                    v.copy_(p.grad)
                p.grad = v
        self._event = None

    def pack(self):
        if self.alias_grads:
            return
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)

    def unpack(self):
        if self.alias_grads:
            return
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)

    def _reduce(self):
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(self.flat, group=self.group)
            self.flat.div_(dist.get_world_size(self.group))

    def all_reduce_mean(self):
        """pack → all-reduce (mean) → (after wait) unpack."""
        self.pack()
        if self.comm_stream is not None:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ready)
                self._reduce()
                self._event = torch.cuda.Event()
                self._event.record(self.comm_stream)
        else:
            self._reduce()

    def wait(self):
        if self._event is not None:
            torch.cuda.current_stream().wait_event(self._event)
            self._event = None
        self.unpack()
