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
    """params: the parameters to train.  duplicated: parameters the reference's optimizer holds TWICE (train.py:198-200
    appends `models["encoder"]` and `models["encoder_mf"]`, one module under shared_encoder / shared_all): torch then
    counts their gradient norm twice, clips them twice and steps them twice per iteration; `reference_duplicates=True` in
    TrainStep reproduces exactly that, the default (each parameter once) is the semantics the reference's authors
    presumably intended.  The learning rate lives on the device (`set_lr`), so a captured step follows a scheduler."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_norm=0.0, distributed=False,
                 duplicated=()):
        dup_ids = {id(p) for p in duplicated}
        seen, first, rest = set(), [], []
        for p in params:
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                (first if id(p) in dup_ids else rest).append(p)
        self.candidates = first + rest          # duplicated parameters lead the arena
        self._dup_ids = dup_ids
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.distributed = distributed
        self.params = None  # fixed after the first backward
        self._skip = None
        self._skip_key = None
        # overlapped all-reduce (enable_overlap): buckets of the gradient arena are gathered and reduced on a communication stream
        # as soon as autograd has produced all of their gradients, while the rest of the backward still runs
        self._overlap = None

    def enable_overlap(self, streams, n_buckets=6):
        """streams: every CUDA stream the backward produces parameter gradients on (the reducer waits for them before it reads a
        bucket).  Takes effect from the second step on (the arena layout is learnt in the first)."""
        if self.distributed and dist.is_initialized() and dist.get_world_size() > 1:
            self._overlap = {"streams": [s for s in streams if s is not None], "n": int(n_buckets), "ready": False}

    def _setup_overlap(self):
        ov = self._overlap
        dev = self.P.device
        ov["comm"] = torch.cuda.Stream(device=dev)
        n = len(self.params)
        target = max(1, self.n // ov["n"])
        bounds, start, acc = [], 0, 0
        for i in range(n):
            acc += (self.sizes[i] + 3) // 4 * 4
            if acc >= target and i + 1 < n:
                bounds.append((start, i + 1))
                start, acc = i + 1, 0
        bounds.append((start, n))
        ov["bounds"] = bounds
        ov["bucket_of"] = [b for b, (lo, hi) in enumerate(bounds) for _ in range(lo, hi)]
        ov["pending"] = [hi - lo for lo, hi in bounds]
        ov["flushed"] = [False] * len(bounds)
        ov["keep"] = []
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(lambda _p, i=i: self._on_grad(i))
        ov["ready"] = True

    def _on_grad(self, i):
        ov = self._overlap
        if ov is None or not ov["ready"] or not ov.get("armed", False):
            return
        b = ov["bucket_of"][i]
        ov["pending"][b] -= 1
        if ov["pending"][b] == 0 and not ov["flushed"][b]:
            self._flush(b)

    def _flush(self, b):
        """gather bucket b's gradients into the arena and all-reduce that slice, on the communication stream"""
        ov = self._overlap
        lo, hi = ov["bounds"][b]
        comm = ov["comm"]
        cur = torch.cuda.current_stream(self.P.device)
        comm.wait_stream(cur)
        for s in ov["streams"]:
            comm.wait_stream(s)
        grads, ptrs = [], []
        for p in self.params[lo:hi]:
            g = p.grad
            if g is not None and (g.dtype != torch.float32 or not g.is_contiguous()):
                g = g.float().contiguous()
            grads.append(g)
            ptrs.append(None if g is None else g.data_ptr())
        n = hi - lo
        a0 = self.offs[lo]
        a1 = self.offs[hi] if hi < len(self.params) else self.n
        with torch.cuda.stream(comm):
            _lib.check(_lib.lib().mvf_gather_grads(self.G.data_ptr(), (ctypes.c_void_p * n)(*ptrs), (ctypes.c_longlong * n)(*self.offs[lo:hi]),
                                                   (ctypes.c_longlong * n)(*self.sizes[lo:hi]), n, comm.cuda_stream), "mvf_gather_grads")
            dist.all_reduce(self.G[a0:a1], op=dist.ReduceOp.SUM)
        ov["keep"].append(grads)    # temporaries stay alive until the step is over
        ov["flushed"][b] = True

    # -- first step: find the parameters that got a gradient, build the arenas, adopt the gradients just computed
    def _build(self):
        self.params = [p for p in self.candidates if p.grad is not None]
        dev = self.params[0].device
        offs, total, n_dup = [], 0, 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4  # every tensor starts 16-byte aligned
            if id(p) in self._dup_ids:
                n_dup = total
        self.n, self.n_dup = total, n_dup
        self.P = torch.zeros(total, device=dev, dtype=torch.float32)
        self.G = torch.zeros(total, device=dev, dtype=torch.float32)
        self.M = torch.zeros(total, device=dev, dtype=torch.float32)
        self.V = torch.zeros(total, device=dev, dtype=torch.float32)
        world = dist.get_world_size() if (self.distributed and dist.is_initialized()) else 1
        self.state = torch.tensor([0.0, 0.0, float(self.lr), 1.0 / world], device=dev, dtype=torch.float32)
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

    def set_lr(self, lr):
        """what a scheduler does to param_groups[...]["lr"] (train.py:289, 668); a device write, visible to graph replays"""
        self.lr = float(lr)
        if self.params is not None:
            self.state[2:3].fill_(self.lr)

    def zero_grad(self):
        """call before backward: with .grad None autograd adopts the tensors it computes instead of accumulating"""
        for p in (self.candidates if self.params is None else self.params):
            p.grad = None
        ov = self._overlap
        if ov is not None and self.params is not None:
            if not ov["ready"]:
                self._setup_overlap()
            ov["pending"] = [hi - lo for lo, hi in ov["bounds"]]
            ov["flushed"] = [False] * len(ov["bounds"])
            ov["keep"] = []
            ov["armed"] = True

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
        self._update_skip(grads)

    def _update_skip(self, grads):
        # parameters without a gradient this step are left untouched by the update (torch skips `grad is None`)
        missing = tuple(i for i, g in enumerate(grads) if g is None)
        if missing != self._skip_key:
            self._skip_key = missing
            if missing:
                mask = torch.zeros(self.n // 4, dtype=torch.uint8)
                for i in missing:
                    mask[self.offs[i] // 4:(self.offs[i] + self.sizes[i] + 3) // 4] = 1
                self._skip = mask.to(self.P.device)
            else:
                self._skip = None

    def step(self):
        if self.params is None:
            self._build()
        ov = self._overlap
        if ov is not None and ov["ready"] and ov.get("armed", False):
            # buckets whose gradients all arrived were reduced during the backward; the rest (a parameter without a gradient this
            # step keeps its bucket open) go now, then this stream joins the communication stream
            for b in range(len(ov["bounds"])):
                if not ov["flushed"][b]:
                    self._flush(b)
            torch.cuda.current_stream(self.P.device).wait_stream(ov["comm"])
            ov["armed"] = False
            self._update_skip([p.grad for p in self.params])
        else:
            self._gather()
            if self.distributed and dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(self.G, op=dist.ReduceOp.SUM)   # the 1 / world scale is applied inside the update kernel
        st = torch.cuda.current_stream(self.P.device).cuda_stream
        from . import conv_tc
        conv_tc.weights_epoch += 1  # the parameters change underneath their tensors' version counters
        _lib.check(_lib.lib().mvf_adamw_step(self.P.data_ptr(), self.G.data_ptr(), self.M.data_ptr(), self.V.data_ptr(), self.n,
                                             self.n_dup, None if self._skip is None else self._skip.data_ptr(),
                                             self.state.data_ptr(), self.ws.data_ptr(), self.ws.numel(), self.betas[0],
                                             self.betas[1], self.eps, self.weight_decay, self.max_norm, st), "mvf_adamw_step")

    @property
    def grad_norm(self):
        return self.state[1]

    # -- checkpointing: the reference saves optimizer.state_dict() in ckpt.pth (train.py:1108-1136)
    def state_dict(self):
        if self.params is None:
            return {"step": 0, "lr": self.lr, "built": False}
        return {"built": True, "step": float(self.state[0]), "lr": self.lr, "sizes": list(self.sizes),
                "exp_avg": self.M.detach().cpu().clone(), "exp_avg_sq": self.V.detach().cpu().clone()}

    def load_state_dict(self, sd):
        self.lr = float(sd["lr"])
        if not sd.get("built", False):
            return
        if self.params is None:
            raise RuntimeError("FlatAdamW.load_state_dict: run one step (or call build_from_grads) before restoring the moments")
        if list(sd["sizes"]) != list(self.sizes):
            raise ValueError("FlatAdamW.load_state_dict: parameter layout differs from the checkpoint's")
        self.M.copy_(sd["exp_avg"])
        self.V.copy_(sd["exp_avg_sq"])
        self.state[0:1].fill_(float(sd["step"]))
        self.state[2:3].fill_(self.lr)
