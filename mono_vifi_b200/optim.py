"""clip_grad_norm_ + AdamW as a handful of kernel launches over flat fp32 arenas (C ABI: mvf_gather_grads +
mvf_adamw_step, csrc/optim.cu).

`FlatAdamW` moves the parameters that actually receive gradients into ONE contiguous buffer (each `p.data` becomes a
view of it, so modules, state_dicts and checkpoints are unaffected) and keeps both Adam moments in two more.  `.grad`
is None between steps, so autograd ADOPTS each gradient tensor it produces (no `grad += new` launch per parameter); one
multi-tensor launch per 128 parameters then gathers them into the flat gradient arena.  Under data parallelism that
arena is all-reduced with a single NCCL call.  Replaces train.py:661-666 (clip_grad_norm_ + optimizer.step) and the
reference's five DistributedDataParallel wrappers (train.py:205-208).  Parameters that never receive a gradient
(torchvision's unused `fc`) are left out, which is what the reference's optimizer does with `grad is None`."""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


class FlatAdamW:
    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_norm=0.0, distributed=False):
        seen, self.candidates = set(), []
        for p in params:
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                self.candidates.append(p)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.distributed = distributed
        self.params = None  # fixed after the first backward

    # -- first step: find the parameters that got a gradient, build the arenas, adopt the gradients just computed
    def _build(self):
        self.params = [p for p in self.candidates if p.grad is not None]
        dev = self.params[0].device
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4  # every tensor starts 16-byte aligned
        self.n = total
        self.P = torch.zeros(total, device=dev, dtype=torch.float32)
        self.G = torch.zeros(total, device=dev, dtype=torch.float32)
        self.M = torch.zeros(total, device=dev, dtype=torch.float32)
        self.V = torch.zeros(total, device=dev, dtype=torch.float32)
        self.state = torch.zeros(2, device=dev, dtype=torch.float32)
        self.ws = torch.empty(_lib.lib().mvf_adamw_workspace_bytes(), device=dev, dtype=torch.uint8)
        self.offs = offs
        self.sizes = [p.numel() for p in self.params]
        n = len(self.params)
        self._c_offs = (ctypes.c_longlong * n)(*offs)
        self._c_sizes = (ctypes.c_longlong * n)(*self.sizes)
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                n = p.numel()
                self.P[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.P[o:o + n].view(p.shape)

    def zero_grad(self):
        """call before backward: with .grad None autograd adopts the tensors it computes instead of accumulating"""
        for p in (self.candidates if self.params is None else self.params):
            p.grad = None

    def _gather(self):
        grads, ptrs = [], []
        for p in self.params:
            g = p.grad
            if g is not None and (g.dtype != torch.float32 or not g.is_contiguous()):
                g = g.float().contiguous()
            grads.append(g)  # keeps temporaries alive until the launch below is enqueued
            ptrs.append(None if g is None else g.data_ptr())
        n = len(ptrs)
        st = torch.cuda.current_stream(self.P.device).cuda_stream
        _lib.check(_lib.lib().mvf_gather_grads(self.G.data_ptr(), (ctypes.c_void_p * n)(*ptrs), self._c_offs, self._c_sizes, n, st),
                   "mvf_gather_grads")

    def step(self):
        if self.params is None:
            self._build()
        self._gather()
        if self.distributed and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.G, op=dist.ReduceOp.SUM)
            self.G.mul_(1.0 / dist.get_world_size())
        st = torch.cuda.current_stream(self.P.device).cuda_stream
        from . import conv_tc
        conv_tc.weights_epoch += 1  # the parameters change underneath their tensors' version counters
        _lib.check(_lib.lib().mvf_adamw_step(self.P.data_ptr(), self.G.data_ptr(), self.M.data_ptr(), self.V.data_ptr(), self.n,
                                             self.state.data_ptr(), self.ws.data_ptr(), self.ws.numel(), self.lr, self.betas[0],
                                             self.betas[1], self.eps, self.weight_decay, self.max_norm, st), "mvf_adamw_step")

    @property
    def grad_norm(self):
        return self.state[1]
