"""Batch-dimension data parallelism: one process per GPU, gradients averaged over NCCL (NVLink 5 / NVSwitch).

The reference wraps each of its 5-6 models in its own DistributedDataParallel (train.py:205-208), i.e. dozens of
25 MB-bucket all-reduces per step plus find_unused_parameters bookkeeping.  Here every parameter's `.grad` is a
view into ONE flat fp32 arena, so a step needs a single all-reduce (~127 MB for the ResNet18 configuration; the
switch reduces it in-fabric when NVLS is available).  Parameters that never receive a gradient (torchvision's unused
`fc`) are detected on the first step and left with `.grad = None`, which is what the reference's
optimizer sees under find_unused_parameters=True.

Works with any torch.distributed backend (the CPU tests use gloo, world_size 2).
"""
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """env:// rendezvous as launched by torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*), train.py:1179-1183."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 0, 1
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        # NCCL writes its version / debug lines to stdout unless told otherwise; stdout is reserved for results
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if not dist.is_initialized():
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, init_method="env://", **kw)
    return rank, local, world


def shard_batch(indices, rank, world):
    """CustomDistributedSampler's strided split (datasets/__init__.py:64-77): indices[rank::world]."""
    return indices[rank::world]


def broadcast_parameters(params, src=0):
    """All ranks start from rank 0's weights (what DDP does at construction)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for p in params:
        dist.broadcast(p.data, src)


def broadcast_module_state(modules, src=0):
    """Parameters AND buffers (BatchNorm running statistics, num_batches_tracked) of every module from rank `src`: what
    DistributedDataParallel does at construction (train.py:208).  With SyncBatchNorm the buffers then stay identical on
    every rank, so a checkpoint does not depend on which rank writes it."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    seen = set()
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            if id(t) not in seen:
                seen.add(id(t))
                dist.broadcast(t.data, src)


class FlatGradAllReduce:
    def __init__(self, params):
        seen, self.params = set(), []
        for p in params:  # the reference lists shared-encoder parameters twice; reduce them once
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                self.params.append(p)
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.arena = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.arena[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.used = None  # learnt on the first step
        self._fired = [False] * len(self.params)
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))
        self.attach()

    def _make_hook(self, i):
        def hook(_p):
            self._fired[i] = True
        return hook

    def attach(self):
        """Point every (used) parameter's .grad at its arena slice and zero the arena.  Call before backward."""
        self.arena.zero_()
        for i, p in enumerate(self.params):
            if self.used is None or self.used[i]:
                p.grad = self.views[i]
            else:
                p.grad = None

    def allreduce_mean(self):
        if self.used is None:
            self.used = list(self._fired)
            for i, p in enumerate(self.params):
                if not self.used[i]:
                    p.grad = None
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.arena, op=dist.ReduceOp.SUM)
            self.arena.mul_(1.0 / dist.get_world_size())
        return self.arena
