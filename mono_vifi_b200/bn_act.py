"""Training-mode BatchNorm2d fused with the residual add and ReLU that follow it in the ResNet / HRNet blocks
(C ABI: mvf_bn_relu_fwd / _bwd, csrc/bn_cl.cu).  `bn_act(bn, x, identity, relu)` keeps nn.BatchNorm2d as the owner of the
parameters and running statistics (state_dict unchanged) and only replaces the arithmetic; anything the kernels do not
cover (eval mode, CPU tensors, channel counts that are not multiples of 4, cumulative-average momentum) takes the plain
torch path `relu(bn(x) + identity)`."""
import contextlib
import os

import torch
import torch.nn.functional as F

from . import _lib

enabled = os.environ.get("MVF_FUSED_BN", "1") != "0"
launches = {"bn_fwd": 0, "bn_bwd": 0}
_ws = {}


def _workspace(dev, n):
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws.get(key)
    if ws is None or ws.numel() < n:
        ws = torch.empty(max(n, 1 << 18), device=dev, dtype=torch.float32)
        _ws[key] = ws
    return ws


def _dense_cl(t):
    B, C, H, W = t.shape
    if t.dtype != torch.float32:
        t = t.float()
    if tuple(t.stride()) != (H * W * C, 1, W * C, C):
        out = torch.empty(B, H, W, C, device=t.device, dtype=torch.float32).permute(0, 3, 1, 2)
        out.copy_(t)
        t = out
    return t


class _BNAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, identity, weight, bias, running_mean, running_var, nbt, eps, momentum, relu):
        x = _dense_cl(x)
        B, C, H, W = x.shape
        P = B * H * W
        if identity is not None:
            identity = _dense_cl(identity)
        y = torch.empty(B, H, W, C, device=x.device, dtype=torch.float32).permute(0, 3, 1, 2)
        fold = 1 if C % 4 == 0 else 2     # C % 4 == 2 (HRNet's 18 channels): two pixels = one row of 2C channels, statistics kept per half
        mean = torch.empty(fold * C, device=x.device, dtype=torch.float32)
        invstd = torch.empty(fold * C, device=x.device, dtype=torch.float32)
        L = _lib.lib()
        ws = _workspace(x.device, L.mvf_bn_workspace_floats(P, C))
        st = torch.cuda.current_stream(x.device).cuda_stream
        launches["bn_fwd"] += 1
        _lib.check(L.mvf_bn_relu_fwd(x.data_ptr(), None if identity is None else identity.data_ptr(), y.data_ptr(), weight.data_ptr(),
                                     bias.data_ptr(), None if running_mean is None else running_mean.data_ptr(),
                                     None if running_var is None else running_var.data_ptr(),
                                     None if nbt is None else nbt.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                     ws.data_ptr(), ws.numel(), P, C, eps, momentum, 1 if relu else 0, st), "mvf_bn_relu_fwd")
        ctx.save_for_backward(x, y, weight, mean, invstd)
        ctx.relu, ctx.has_identity = relu, identity is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, y, weight, mean, invstd = ctx.saved_tensors
        B, C, H, W = x.shape
        P = B * H * W
        gy = _dense_cl(gy)
        gx = torch.empty(B, H, W, C, device=x.device, dtype=torch.float32).permute(0, 3, 1, 2)
        gid = torch.empty(B, H, W, C, device=x.device, dtype=torch.float32).permute(0, 3, 1, 2) if (
            ctx.has_identity and ctx.needs_input_grad[1]) else None
        dgamma = torch.empty(C, device=x.device, dtype=torch.float32)
        dbeta = torch.empty(C, device=x.device, dtype=torch.float32)
        L = _lib.lib()
        ws = _workspace(x.device, L.mvf_bn_workspace_floats(P, C))
        st = torch.cuda.current_stream(x.device).cuda_stream
        launches["bn_bwd"] += 1
        _lib.check(L.mvf_bn_relu_bwd(x.data_ptr(), gy.data_ptr(), y.data_ptr(), weight.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                     gx.data_ptr(), None if gid is None else gid.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                                     ws.data_ptr(), ws.numel(), P, C, 1 if ctx.relu else 0, st), "mvf_bn_relu_bwd")
        return gx, gid, dgamma, dbeta, None, None, None, None, None, None


# ---- SyncBatchNorm (train.py:205-208) -----------------------------------------------------------------------------------
# A BatchNorm the reference's `nn.SyncBatchNorm.convert_sync_batchnorm` has converted (or TrainStep(distributed=True) did
# the same way) normalises with the statistics of ALL ranks.  Same kernels, split around one all-reduce of a
# [2C + 1] float64 vector per call (sum, sum of squares / sum g, sum g*xhat, pixel count) on the module's process group.
# Collectives issued from different CUDA streams (the concurrent pose / depth branches) use one NCCL communicator per
# stream (`group_for_stream`) so that they neither serialise nor interleave differently across ranks.
_groups = {}


def group_for_stream(base_group, stream):
    import torch.distributed as dist
    key = (id(base_group), stream.cuda_stream)
    g = _groups.get(key)
    if g is None:
        first = not any(k[0] == id(base_group) for k in _groups)
        # the first stream that asks keeps the module's own group; further streams get their own communicator.  Every
        # rank runs the same program, so every rank creates the groups in the same order (new_group is collective).
        g = base_group if first else dist.new_group(ranks=dist.get_process_group_ranks(base_group) if base_group is not None else None)
        _groups[key] = g
    return g


def _exchange(sums, group, stream):
    """cross-rank SUM of the statistics vector: one single-CTA kernel over NVLink peer memory (peer.py) when the node offers
    symmetric memory, else an NCCL all-reduce on the stream's own communicator"""
    import torch.distributed as dist
    from . import peer
    px = peer.get(group, sums.device)
    if px is not None:
        px.allreduce_(sums, stream)
    else:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group_for_stream(group, stream))


def _sync_world(bn):
    import torch.distributed as dist
    if not isinstance(bn, torch.nn.SyncBatchNorm) or not dist.is_available() or not dist.is_initialized():
        return None, 1
    group = bn.process_group if bn.process_group is not None else dist.group.WORLD
    return group, dist.get_world_size(group)


class _BNActSync(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, identity, weight, bias, running_mean, running_var, nbt, eps, momentum, relu, group):
        import torch.distributed as dist
        x = _dense_cl(x)
        B, C, H, W = x.shape
        P = B * H * W
        if identity is not None:
            identity = _dense_cl(identity)
        y = torch.empty(B, H, W, C, device=x.device, dtype=torch.float32).permute(0, 3, 1, 2)
        mean = torch.empty(C, device=x.device, dtype=torch.float32)
        invstd = torch.empty(C, device=x.device, dtype=torch.float32)
        sums = torch.empty(2 * C + 1, device=x.device, dtype=torch.float64)
        L = _lib.lib()
        ws = _workspace(x.device, L.mvf_bn_workspace_floats(P, C))
        cur = torch.cuda.current_stream(x.device)
        st = cur.cuda_stream
        launches["bn_fwd"] += 1
        launches["bn_sync"] = launches.get("bn_sync", 0) + 1
        _lib.check(L.mvf_bn_sync_stats_fwd(x.data_ptr(), sums.data_ptr(), ws.data_ptr(), ws.numel(), P, C, st), "mvf_bn_sync_stats_fwd")
        _exchange(sums, group, cur)
        _lib.check(L.mvf_bn_sync_apply_fwd(x.data_ptr(), None if identity is None else identity.data_ptr(), y.data_ptr(),
                                           weight.data_ptr(), bias.data_ptr(), None if running_mean is None else running_mean.data_ptr(),
                                           None if running_var is None else running_var.data_ptr(),
                                           None if nbt is None else nbt.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                           sums.data_ptr(), P, C, eps, momentum, 1 if relu else 0, st), "mvf_bn_sync_apply_fwd")
        ctx.save_for_backward(x, y, weight, mean, invstd)
        ctx.relu, ctx.has_identity, ctx.group = relu, identity is not None, group
        return y

    @staticmethod
    def backward(ctx, gy):
        import torch.distributed as dist
        x, y, weight, mean, invstd = ctx.saved_tensors
        B, C, H, W = x.shape
        P = B * H * W
        gy = _dense_cl(gy)
        gx = torch.empty(B, H, W, C, device=x.device, dtype=torch.float32).permute(0, 3, 1, 2)
        gid = torch.empty(B, H, W, C, device=x.device, dtype=torch.float32).permute(0, 3, 1, 2) if (
            ctx.has_identity and ctx.needs_input_grad[1]) else None
        dgamma = torch.empty(C, device=x.device, dtype=torch.float32)
        dbeta = torch.empty(C, device=x.device, dtype=torch.float32)
        sums = torch.empty(2 * C + 1, device=x.device, dtype=torch.float64)
        scratch = torch.empty(2 * C + 4, device=x.device, dtype=torch.float32)
        L = _lib.lib()
        ws = _workspace(x.device, L.mvf_bn_workspace_floats(P, C))
        cur = torch.cuda.current_stream(x.device)
        st = cur.cuda_stream
        launches["bn_bwd"] += 1
        _lib.check(L.mvf_bn_sync_stats_bwd(x.data_ptr(), gy.data_ptr(), y.data_ptr(), mean.data_ptr(), invstd.data_ptr(), sums.data_ptr(),
                                           dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), ws.numel(), P, C, 1 if ctx.relu else 0, st),
                   "mvf_bn_sync_stats_bwd")
        _exchange(sums, ctx.group, cur)
        _lib.check(L.mvf_bn_sync_apply_bwd(x.data_ptr(), gy.data_ptr(), y.data_ptr(), weight.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                           gx.data_ptr(), None if gid is None else gid.data_ptr(), sums.data_ptr(), scratch.data_ptr(),
                                           scratch.numel(), P, C, 1 if ctx.relu else 0, st), "mvf_bn_sync_apply_bwd")
        return gx, gid, dgamma, dbeta, None, None, None, None, None, None, None


def usable(bn, x):
    if not (enabled and bn.training and x.is_cuda and x.dim() == 4 and x.shape[1] <= 1024 and bn.affine and bn.momentum is not None and
            x.dtype == torch.float32):
        return False
    C = x.shape[1]
    if C % 4 == 0:
        return True
    # C % 4 == 2: the single-process kernels fold two pixels into one row (the cross-rank SyncBatchNorm split does not)
    return C % 2 == 0 and (x.shape[0] * x.shape[2] * x.shape[3]) % 2 == 0 and _sync_world(bn)[1] == 1


# ---- deferred running statistics ------------------------------------------------------------------------------------
# When the same BatchNorm modules are applied twice CONCURRENTLY (the two pose passes on two streams), the running
# statistics must still see the two updates in program order.  Inside `with deferred_running_stats() as entries:` a call
# normalises as usual but writes its batch mean / unbiased variance to a private buffer (momentum 1 into zeros) instead
# of touching the module; apply_deferred(entries), called later on a stream that has joined both passes, replays the
# momentum updates in order with three multi-tensor launches.
_deferred = None


@contextlib.contextmanager
def deferred_running_stats():
    global _deferred
    prev, _deferred = _deferred, []
    try:
        yield _deferred
    finally:
        _deferred = prev


def apply_deferred(entries):
    if not entries:
        return
    with torch.no_grad():
        groups = {}
        for bn, tmp, m in entries:
            g = groups.setdefault(m, ([], []))
            C = bn.running_mean.numel()
            g[0].extend([bn.running_mean, bn.running_var])
            g[1].extend([tmp[:C], tmp[C:]])
        for m, (stats, batch) in groups.items():
            torch._foreach_mul_(stats, 1.0 - m)
            torch._foreach_add_(stats, batch, alpha=m)
        counters = [bn.num_batches_tracked for bn, _, _ in entries if bn.num_batches_tracked is not None]
        if counters:
            torch._foreach_add_(counters, 1)
    entries.clear()


def bn_act(bn, x, identity=None, relu=True):
    """relu(bn(x) + identity) with nn.BatchNorm2d `bn` (its parameters / buffers are used and updated in place)."""
    if usable(bn, x):
        rm, rv, nbt = (bn.running_mean, bn.running_var, bn.num_batches_tracked) if bn.track_running_stats else (None, None, None)
        if rm is not None and _deferred is not None:
            C = rm.numel()
            tmp = torch.zeros(2 * C, device=x.device, dtype=torch.float32)
            _deferred.append((bn, tmp, float(bn.momentum)))
            group, world = _sync_world(bn)
            if world > 1:
                return _BNActSync.apply(x, identity, bn.weight, bn.bias, tmp[:C], tmp[C:], None, float(bn.eps), 1.0, bool(relu), group)
            return _BNAct.apply(x, identity, bn.weight, bn.bias, tmp[:C], tmp[C:], None, float(bn.eps), 1.0, bool(relu))
        if nbt is not None and nbt.dtype != torch.int64:
            nbt.add_(1)
            nbt = None
        group, world = _sync_world(bn)
        if world > 1:
            return _BNActSync.apply(x, identity, bn.weight, bn.bias, rm, rv, nbt, float(bn.eps), float(bn.momentum), bool(relu), group)
        return _BNAct.apply(x, identity, bn.weight, bn.bias, rm, rv, nbt, float(bn.eps), float(bn.momentum), bool(relu))
    if (enabled and not bn.training and x.is_cuda and x.dim() == 4 and x.shape[1] % 4 == 0 and bn.affine and bn.track_running_stats and
            x.dtype == torch.float32 and not (torch.is_grad_enabled() and (x.requires_grad or bn.weight.requires_grad))):
        # inference (evaluation loop, train.py:419-483): running statistics, one pass (mvf_bn_eval_fwd)
        x = _dense_cl(x)
        B, C, H, W = x.shape
        if identity is not None:
            identity = _dense_cl(identity)
        y = torch.empty(B, H, W, C, device=x.device, dtype=torch.float32).permute(0, 3, 1, 2)
        launches["bn_eval"] = launches.get("bn_eval", 0) + 1
        _lib.check(_lib.lib().mvf_bn_eval_fwd(x.data_ptr(), None if identity is None else identity.data_ptr(), y.data_ptr(),
                                              bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                                              bn.running_var.data_ptr(), B * H * W, C, float(bn.eps), 1 if relu else 0,
                                              torch.cuda.current_stream(x.device).cuda_stream), "mvf_bn_eval_fwd")
        return y
    y = bn(x)
    if identity is not None:
        y = y + identity
    return F.relu(y) if relu else y
