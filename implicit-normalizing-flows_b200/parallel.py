"""Batch-sharded data parallelism: one process per GPU, one flat-bucket all-reduce of the
gradients per step.  Replaces the reference's torch.nn.DataParallel
(train_img.py:203-204,820).

The hot path itself has no exchange step — solves, probes and log-dets are per-sample
independent — so the only collective is the gradient reduction.  Works with any
torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist

__all__ = ['shard_batch', 'broadcast_module', 'FlatGradBucket', 'allreduce_mean_scalar', 'sink_grads']


def shard_batch(x, rank=None, world_size=None):
    """Contiguous split of the global batch along dim 0 (the same split DataParallel.scatter makes)."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    B = x.shape[0]
    per = (B + world_size - 1) // world_size
    return x[rank * per:min(B, (rank + 1) * per)]


def broadcast_module(module, src=0):
    """Rank `src` initialises (data-dependent ActNorm init, lazy u/v shaping); everyone else
    receives parameters and buffers."""
    with torch.no_grad():
        tensors = list(module.parameters()) + list(module.buffers())
        for t in tensors:
            dist.broadcast(t.data, src=src)
        # the collective writes through the raw storage: bump the version counters so that every host cache keyed
        # on them (effective weights, sigma, d sigma / dW, roulette rates) is rebuilt from the received values
        torch.autograd.graph.increment_version(tensors)
        for m in module.modules():           # python mirrors of buffers (lazy conv u / v shapes, ActNorm init flag)
            if hasattr(m, '_hw'):
                m._hw = None
            if hasattr(m, '_init_known'):
                m._init_known = None


class FlatGradBucket(object):
    """Views every parameter's .grad into one flat fp32 buffer so that the step's gradient
    exchange is a single all-reduce (21.9 MB for the CIFAR config: latency-, not bandwidth-bound
    on NVLink 5, so one bucket is the right granularity)."""

    ALIGN = 64      # floats

    def __init__(self, params, direct=True):
        """direct: the graph-free backward sweeps of this package add their finished parameter gradients straight
        into the bucket (sink_grads) instead of handing them to autograd's per-parameter AccumulateGrad nodes."""
        self.params = [p for p in params if p.requires_grad]
        self.direct = direct
        self._armed = False
        # every tensor starts on a 256-byte boundary (the kernels read biases / weights with vector loads);
        # the padding stays zero, so norms and the all-reduce are unaffected
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        # total length a multiple of 1024 floats: whole-buffer reductions (the clip norm) split evenly over 256 CTAs
        off = (off + 1023) // 1024 * 1024
        dev = self.params[0].device if self.params else torch.device('cpu')
        self.flat = torch.zeros(off, device=dev, dtype=torch.float32)
        self.views = []
        self.had_grad = [True] * len(self.params)      # of the latest gather_strays()
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            v = self.flat[o:o + p.numel()].view_as(p)
            p.grad = v
            self.views.append(v)
            if direct:
                p._impflow_grad_sink = (self, i)
        self._touched = [False] * len(self.params)

    def zero(self):
        """Start of a step.  The .grad slots are emptied rather than pointed at the (zeroed) views: autograd then
        keeps each parameter's gradient tensor as it arrives (no accumulate kernel per parameter) and
        gather_strays() moves all of them into the flat buffer with one multi-tensor copy."""
        self.flat.zero_()
        for p in self.params:
            p.grad = None
        self._touched = [False] * len(self.params)
        self._armed = self.direct       # between zero() and the gather, finished gradients may be added directly

    def sink(self, idxs, grads):
        """views[i] += grads[i] for all i in one multi-tensor launch (called by sink_grads)."""
        with torch.no_grad():
            torch._foreach_add_([self.views[i].view(-1) for i in idxs], [g.detach().reshape(-1) for g in grads])
        for i in idxs:
            self._touched[i] = True

    def gather_strays(self):
        """Copy the gradients autograd produced outside the bucket into it (one multi-tensor launch) and point
        every .grad at its view; parameters without a gradient keep the zeros."""
        if all(p.grad is v for p, v in zip(self.params, self.views)):
            return      # nothing outside the bucket (e.g. the optimiser's call after allreduce_mean's)
        dst, src, add_dst, add_src, empty = [], [], [], [], []
        touched = self._touched
        self.had_grad = [p.grad is not None or t for p, t in zip(self.params, touched)]
        for p, v, t in zip(self.params, self.views, touched):
            g = p.grad
            if g is None:
                if not t:
                    empty.append(v)    # stays zero also when the caller used zero_grad() instead of zero()
            elif g.data_ptr() != v.data_ptr():
                # flat views on both sides: the multi-tensor fast path needs equal strides (size-1 dimensions
                # of conv weights can carry different ones), else it degrades to one cudaMemcpy per tensor
                if t:                  # part of the gradient was added directly, the rest came through autograd
                    add_dst.append(v.view(-1))
                    add_src.append(g.detach().reshape(-1))
                else:
                    dst.append(v.view(-1))
                    src.append(g.detach().reshape(-1))
            p.grad = v
        self._armed = False
        self._touched = [False] * len(self.params)
        with torch.no_grad():
            if empty:
                torch._foreach_zero_(empty)
            if dst:
                torch._foreach_copy_(dst, src)
            if add_dst:
                torch._foreach_add_(add_dst, add_src)

    def allreduce_mean(self, group=None):
        self.gather_strays()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))
        return self.flat


def sink_grads(params, grads):
    """Finished parameter gradients of a graph-free backward sweep -> straight into the flat bucket (one multi-tensor
    add per bucket) where the parameter has one that is armed (between FlatGradBucket.zero() and the gather); the
    entries that were taken come back as None, so autograd has nothing to accumulate for them.  Without a bucket
    the list is returned unchanged."""
    out, per_bucket = list(grads), {}
    for i, (p, g) in enumerate(zip(params, grads)):
        if g is None:
            continue
        s = getattr(p, '_impflow_grad_sink', None)
        if s is None or not s[0]._armed or s[0].params[s[1]] is not p:
            continue
        idxs, gs = per_bucket.setdefault(id(s[0]), (s[0], [], []))[1:]
        idxs.append(s[1])
        gs.append(g)
        out[i] = None
    for bucket, idxs, gs in per_bucket.values():
        bucket.sink(idxs, gs)
    return out


def allreduce_mean_scalar(t, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(dist.get_world_size(group))
    return t
