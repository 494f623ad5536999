"""Training-step orchestration of the hot path (host side).

Mirrors the part of the reference's `Trainer` that BASELINE.json's north_star names -- `predict_poses`
(train.py:943-954), the single-frame loss group of `process_batch` (train.py:728-729, 736, 739, 747-750),
backward, gradient clipping and AdamW (train.py:659-666) -- on top of the drop-in networks and the fused
photometric-loss kernel.  Data-parallel training shards the batch across ranks; gradients are averaged with ONE
NCCL all-reduce over a flat fp32 arena (mono_vifi_b200/ddp.py) instead of the reference's five DDP wrappers.
"""
import os

import torch

from . import layers as L
from . import networks as N
from . import affine as A
from . import bn_act
from .ddp import FlatGradAllReduce
from .fused import fused_photometric_loss


class Options:
    """The hot-path subset of options.py (defaults: options.py:68-209)."""

    def __init__(self, **kw):
        self.height, self.width, self.batch_size = 192, 640, 12
        self.min_depth, self.max_depth = 0.1, 100.0
        self.disparity_smoothness = 1e-3
        self.no_ssim = self.avg_reprojection = self.disable_automasking = False
        self.learning_rate, self.weight_decay, self.clip_grad = 1e-4, 0.01, 5.0
        self.num_layers = 18
        self.num_scales = 1
        self.backbone = "ResNet18"              # ResNet18 | ResNet50 | DHRNet | LiteMono   (options.py:187-190)
        self.fuse_model_type = "shared_encoder"  # shared_encoder | separate_all | shared_all (options.py:196-200)
        # IFRNet size of the TRAINING step.  The reference hard-codes IFRNet("large") there (train.py:210) -- its
        # --vfi_scale (options.py:191-195) only picks the test-time model -- so "large" is the default; BASELINE.json
        # configs[3] names IFRNet_S for the Lite-Mono configuration and sets "small" explicitly.
        self.vfi_scale = "large"
        # True: reproduce the reference's optimizer quirk (train.py:198-200): parameters of a module that sits under two
        # keys of self.models (encoder / encoder_mf, and depth / depth_mf under shared_all) are held twice by AdamW and
        # by clip_grad_norm_ -- norm counted twice, clipped twice, stepped twice per iteration.  False: each once.
        self.reference_duplicates = False
        # data-parallel runs: BatchNorm statistics over the batches of all ranks, as the reference's
        # nn.SyncBatchNorm.convert_sync_batchnorm (train.py:205-208) makes them.  False = per-GPU statistics (a deviation
        # from the reference, kept for A/B timing of what the exchange costs).
        self.sync_bn = True
        self.lamda = 0.2                        # weight of the depth-consistency loss (options.py:92-95)
        self.multi_frame = False                # False: the single-frame slice (BASELINE configs[1]); True: full process_batch
        self.tie_break_noise = True  # train.py:1023: torch.randn * 1e-5 on the identity terms
        self.use_affine = False      # the affine-augmentation branch of process_batch (train.py:815-883, options.py:96-99)
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError("unknown option %r" % k)
            setattr(self, k, v)


def build_models(opt, device):
    """train.py:142-190 (weights_init=scratch): depth encoder / decoder of the chosen backbone, pose networks and, for the
    multi-frame branch, the second decoder (or encoder + decoder) and the fusion module."""
    import copy
    models = {}
    if opt.backbone in ("ResNet18", "ResNet50"):
        models["encoder"] = N.monodepth2.DepthEncoder(18 if opt.backbone == "ResNet18" else 50, False)
        models["depth"] = N.monodepth2.DepthDecoder(models["encoder"].num_ch_enc, range(opt.num_scales))
    elif opt.backbone == "DHRNet":
        models["encoder"] = N.DHRNet.DepthEncoder(18, False)
        models["depth"] = N.DHRNet.DepthDecoder(models["encoder"].num_ch_enc, range(opt.num_scales))
    elif opt.backbone == "LiteMono":
        models["encoder"] = N.LiteMono.DepthEncoder(model="lite-mono", drop_path_rate=0.2, width=opt.width, height=opt.height)
        models["depth"] = N.LiteMono.DepthDecoder(models["encoder"].num_ch_enc, range(opt.num_scales))
    else:
        raise ValueError("unknown backbone %r" % (opt.backbone,))
    if opt.multi_frame:
        if opt.fuse_model_type == "shared_all":
            models["encoder_mf"], models["depth_mf"] = models["encoder"], models["depth"]
        elif opt.fuse_model_type == "shared_encoder":
            models["encoder_mf"], models["depth_mf"] = models["encoder"], copy.deepcopy(models["depth"])
        elif opt.fuse_model_type == "separate_all":
            models["encoder_mf"], models["depth_mf"] = copy.deepcopy(models["encoder"]), copy.deepcopy(models["depth"])
        else:
            raise ValueError("unknown fuse_model_type %r" % (opt.fuse_model_type,))
        models["fusion_module"] = N.FusionModule(opt, models["encoder_mf"].num_ch_enc)
    models["pose_encoder"] = N.posenet.ResnetEncoder(opt.num_layers, False, num_input_images=2)
    models["pose"] = N.posenet.PoseDecoder(models["pose_encoder"].num_ch_enc, num_input_features=1,
                                           num_frames_to_predict_for=2)
    for m in models.values():
        m.to(device)
    return models


def predict_poses(models, img_0, img_1):
    """train.py:943-954"""
    feats = [models["pose_encoder"](torch.cat([img_0, img_1], 1))]
    axisangle, translation = models["pose"](feats)
    pose = L.transformation_from_parameters(axisangle[:, 0], translation[:, 0], invert=False)
    pose_inv = L.transformation_from_parameters(axisangle[:, 0], translation[:, 0], invert=True)
    return pose, pose_inv


def loss_group(opt, disp, img_tgt, T0, T1, img_src0, img_src1, K, inv_K, mask_rec=None, noise=None):
    """generate_images_pred x2 + compute_losses_base (train.py:747-750) as ONE fused kernel launch."""
    P0 = L.matmul_KT(K, T0)[:, :3, :]
    P1 = L.matmul_KT(K, T1)[:, :3, :]
    if noise is None and opt.tie_break_noise and not opt.disable_automasking:
        B, _, H, W = disp.shape
        noise = torch.randn(B, 1 if opt.avg_reprojection else 2, H, W, device=disp.device)
    return fused_photometric_loss(disp, img_tgt, img_src0, img_src1, inv_K, P0, P1, noise, mask_rec, opt.min_depth,
                                  opt.max_depth, opt.disparity_smoothness, opt.no_ssim, opt.avg_reprojection,
                                  opt.disable_automasking)


class _Fork:
    """`with _Fork(side):` runs its body on the stream `side`, ordered after everything already queued on the current
    stream; join() makes the current stream wait for it and tells the allocator that the given tensors (allocated on
    `side`) are now used on the current stream.  With side None the body runs inline."""

    def __init__(self, side):
        self.side = side
        self.cur = torch.cuda.current_stream(side.device) if side is not None else None
        self.ctx = None

    def __enter__(self):
        if self.side is not None:
            self.side.wait_stream(self.cur)
            self.ctx = torch.cuda.stream(self.side)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False

    def join(self, *tensors):
        if self.side is not None:
            self.cur.wait_stream(self.side)
            for t in tensors:
                t.record_stream(self.cur)


def single_frame_losses(models, inputs, opt, side=None, side2=None):
    """The single-frame slice of process_batch: train.py:728-729 (poses), 736 + 739 (depth), 747-750 (loss).
    The pose branch and the depth branch are independent until the loss: with `side` (a CUDA stream) the two pose
    passes run there while the depth network runs on the current stream, so their short kernels (1-3 tiles per SM, many
    with fewer CTAs than SMs) fill each other's idle SMs; autograd replays the same split in the backward."""
    img_n1, img_0, img_p1 = inputs[("color", -1, 0)], inputs[("color", 0, 0)], inputs[("color", 1, 0)]
    K, inv_K = inputs[("K", 0)], inputs[("inv_K", 0)]
    with _Fork(side) as fork:
        pose_n1_0, pose_0_n1 = predict_poses(models, inputs[("color_aug", -1, 0)], inputs[("color_aug", 0, 0)])
    # the second pose pass shares its weights with the first; on its own stream (side2) it defers its BatchNorm
    # running-statistics updates, which are replayed in program order once both passes have been joined
    with _Fork(side2 if side2 is not None else side) as fork2:
        if side2 is not None:
            with bn_act.deferred_running_stats() as deferred:
                pose_0_p1, pose_p1_0 = predict_poses(models, inputs[("color_aug", 0, 0)], inputs[("color_aug", 1, 0)])
        else:
            deferred = None
            pose_0_p1, pose_p1_0 = predict_poses(models, inputs[("color_aug", 0, 0)], inputs[("color_aug", 1, 0)])
    disp_0 = models["depth"](models["encoder"](inputs[("color_aug", 0, 0)]))[("disp", 0)]
    fork.join(pose_0_n1)
    fork2.join(pose_0_p1)
    bn_act.apply_deferred(deferred)
    loss, auto_mask = loss_group(opt, disp_0, img_0, pose_0_n1, pose_0_p1, img_n1, img_p1, K, inv_K)
    return {"loss": loss, "loss_base": loss, "disp": disp_0, "auto_mask": auto_mask}


def multi_frame_losses(models, vfi, inputs, opt):
    """The full process_batch of the reference (train.py:698-885; the affine branch with opt.use_affine): three VFI passes
    (frozen IFRNet), six pose passes, single-frame depth of the target and of the two synthesized frames, fused
    multi-frame depth of the same three, six photometric loss groups (each ONE fused kernel launch) and three
    scale-invariant log depth-consistency terms.  loss = loss_base + lamda * loss_dc."""
    img_n1, img_0, img_p1 = inputs[("color", -1, 0)], inputs[("color", 0, 0)], inputs[("color", 1, 0)]
    aug_n1, aug_0, aug_p1 = inputs[("color_aug", -1, 0)], inputs[("color_aug", 0, 0)], inputs[("color_aug", 1, 0)]
    K, inv_K = inputs[("K", 0)], inputs[("inv_K", 0)]
    embt = torch.full((img_0.shape[0], 1, 1, 1), 0.5, device=img_0.device)
    with torch.no_grad():
        img_nt, flow_nt_n1, flow_nt_0, mask_nt = vfi(img_n1, img_0, embt)
        img_pt, flow_pt_0, flow_pt_p1, mask_pt = vfi(img_0, img_p1, embt)
        flow_0_n1, flow_0_p1, mask_01 = vfi(img_n1, img_p1, embt, onlyFlow=True)
    _, pose_0_n1 = predict_poses(models, aug_n1, aug_0)
    pose_0_p1, _ = predict_poses(models, aug_0, aug_p1)
    _, pose_nt_n1 = predict_poses(models, img_n1, img_nt)
    pose_nt_p1, _ = predict_poses(models, img_nt, img_p1)
    _, pose_pt_n1 = predict_poses(models, img_n1, img_pt)
    pose_pt_p1, _ = predict_poses(models, img_pt, img_p1)

    enc, dec = models["encoder"], models["depth"]
    keep = lambda feats: list(feats)  # the reference's encoders overwrite self.features on every call
    feats_0, feats_nt, feats_pt = keep(enc(aug_0)), keep(enc(img_nt)), keep(enc(img_pt))
    disp_0, disp_nt, disp_pt = dec(feats_0)[("disp", 0)], dec(feats_nt)[("disp", 0)], dec(feats_pt)[("disp", 0)]
    loss_base = 0.0
    group = lambda disp, tgt, T_n1, T_p1: loss_group(opt, disp, tgt, T_n1, T_p1, img_n1, img_p1, K, inv_K)[0]
    loss_base = group(disp_0, img_0, pose_0_n1, pose_0_p1) + group(disp_pt, img_pt, pose_pt_n1, pose_pt_p1) + \
        group(disp_nt, img_nt, pose_nt_n1, pose_nt_p1)

    enc_mf, dec_mf, fuse = models["encoder_mf"], models["depth_mf"], models["fusion_module"]
    if opt.fuse_model_type == "separate_all":
        feats_0, feats_nt, feats_pt = keep(enc_mf(aug_0)), keep(enc_mf(img_nt)), keep(enc_mf(img_pt))
    feats_n1, feats_p1 = keep(enc_mf(aug_n1)), keep(enc_mf(aug_p1))
    disp_0_f = dec_mf(fuse([feats_n1, feats_0, feats_p1], [flow_0_n1, flow_0_p1], mask_01))[("disp", 0)]
    disp_nt_f = dec_mf(fuse([feats_n1, feats_nt, feats_0], [flow_nt_n1, flow_nt_0], mask_nt))[("disp", 0)]
    disp_pt_f = dec_mf(fuse([feats_0, feats_pt, feats_p1], [flow_pt_0, flow_pt_p1], mask_pt))[("disp", 0)]
    loss_base = loss_base + group(disp_0_f, img_0, pose_0_n1, pose_0_p1) + group(disp_nt_f, img_nt, pose_nt_n1, pose_nt_p1) + \
        group(disp_pt_f, img_pt, pose_pt_n1, pose_pt_p1)
    depth = lambda d: L.disp_to_depth(d, opt.min_depth, opt.max_depth)[1]
    loss_dc = L.si_log_depth_loss(depth(disp_0), depth(disp_0_f)) + L.si_log_depth_loss(depth(disp_nt), depth(disp_nt_f)) + \
        L.si_log_depth_loss(depth(disp_pt), depth(disp_pt_f))
    if opt.use_affine:
        # train.py:815-883: the single-frame network on the rotated / cropped / rescaled view of each of the three frames;
        # poses conjugated by Rc, photometric loss masked by valid_mask_rec, and two scale-aware depth-consistency terms
        # against the un-augmented single-frame and fused depths.  Batched transforms: affine.py (no .item() syncs).
        Rc, angle, box, ratio = inputs["Rc"], inputs["angle"], inputs["box"], inputs["ratio_local"]
        mask_rec, mask_cons = inputs["valid_mask_rec"], inputs["valid_mask_cons"]
        aff_n1, aff_0, aff_p1 = inputs[("color_affine", -1, 0)], inputs[("color_affine", 0, 0)], inputs[("color_affine", 1, 0)]
        conj = lambda T: A.conjugate_pose(T, Rc)

        def affine_terms(net_in, tgt, T_n1, T_p1, disp_single, disp_fused):
            disp_a = dec(keep(enc(net_in)))[("disp", 0)]
            lb = loss_group(opt, disp_a, tgt, conj(T_n1), conj(T_p1), aff_n1, aff_p1, K, inv_K, mask_rec=mask_rec)[0]
            restore = A.depth_restore(depth(disp_a), angle, box, ratio)
            return lb, (L.si_log_depth_loss(restore, depth(disp_fused), mask_cons) +
                        L.si_log_depth_loss(restore, depth(disp_single), mask_cons))

        img_nt_a, img_pt_a = A.affine_transform(img_nt, angle, box), A.affine_transform(img_pt, angle, box)
        for net_in, tgt, T_n1, T_p1, d_s, d_f in ((inputs[("color_affine_aug", 0, 0)], aff_0, pose_0_n1, pose_0_p1, disp_0, disp_0_f),
                                                  (img_nt_a, img_nt_a, pose_nt_n1, pose_nt_p1, disp_nt, disp_nt_f),
                                                  (img_pt_a, img_pt_a, pose_pt_n1, pose_pt_p1, disp_pt, disp_pt_f)):
            lb, ldc = affine_terms(net_in, tgt, T_n1, T_p1, d_s, d_f)
            loss_base, loss_dc = loss_base + lb, loss_dc + ldc
    return {"loss": loss_base + opt.lamda * loss_dc, "loss_base": loss_base, "loss_dc": loss_dc, "disp": disp_0,
            "disp_fuse": disp_0_f}


class TrainStep:
    """zero_grad -> forward -> backward -> (all-reduce) -> clip -> AdamW  (train.py:656-666)."""

    def __init__(self, opt, device, models=None, distributed=False, capturable=False, fused_optimizer=None):
        self.opt, self.device = opt, device
        self.models = models if models is not None else build_models(opt, device)
        if distributed and device.type == "cuda" and getattr(opt, "sync_bn", True):
            import torch.distributed as dist
            if dist.is_initialized() and dist.get_world_size() > 1:
                # train.py:205-208; converts children in place, so modules registered under two keys stay one module
                for k in list(self.models):
                    self.models[k] = torch.nn.SyncBatchNorm.convert_sync_batchnorm(self.models[k])
        # the frozen VFI network of the multi-frame branch (train.py:210-216); not trained, not in the optimizer
        self.vfi = N.IFRNet(opt.vfi_scale).to(device).eval() if opt.multi_frame else None
        # the reference appends every model's parameters (train.py:198-200); a module under two keys is listed twice
        seen, self.params, self.duplicated = {}, [], []
        for m in self.models.values():
            for p in m.parameters():
                if id(p) not in seen:
                    seen[id(p)] = 1
                    self.params.append(p)
                else:
                    seen[id(p)] += 1
                    if seen[id(p)] == 2:
                        self.duplicated.append(p)
        if not getattr(opt, "reference_duplicates", False):
            self.duplicated = []
        # default on CUDA: clip + AdamW (+ the gradient all-reduce) over flat arenas, two kernel launches per step
        # (optim.FlatAdamW); otherwise torch.optim.AdamW (capturable: step counters on the device for CUDA graphs)
        if fused_optimizer is None:
            fused_optimizer = device.type == "cuda"
        self.flat = None
        self.optimizer = None
        self.reducer = None
        if fused_optimizer:
            from .optim import FlatAdamW
            self.flat = FlatAdamW(self.params, lr=opt.learning_rate, weight_decay=opt.weight_decay,
                                  max_norm=float(opt.clip_grad or 0.0), distributed=distributed, duplicated=self.duplicated)
        else:
            # (torch keeps a duplicated list entry: it warns and steps it twice, which is the reference's behaviour)
            self.optimizer = torch.optim.AdamW(self.params + self.duplicated, lr=opt.learning_rate, weight_decay=opt.weight_decay,
                                               capturable=bool(capturable and device.type == "cuda"))
            self.reducer = FlatGradAllReduce(self.params) if distributed else None
        # All training work runs on one dedicated (non-default) stream.  autograd binds each parameter's gradient
        # accumulator to the stream of its first use; binding them to the legacy default stream would make the step
        # impossible to record into a CUDA graph later (GraphedTrainStep).
        # The depth branch + loss + optimiser (main stream) is the longest chain of the step; with MVF_STREAM_PRIORITY=1 it
        # gets a higher scheduling priority than the pose streams and the weight-gradient companions that share the SMs.
        prio = -1 if os.environ.get("MVF_STREAM_PRIORITY", "0") == "1" else 0
        self.stream = torch.cuda.Stream(device=device, priority=prio) if device.type == "cuda" else None
        # streams for the two pose passes of the single-frame step (MVF_SIDE_STREAM=1: both on one, 0: everything serial)
        self.side = (torch.cuda.Stream(device=device)
                     if device.type == "cuda" and os.environ.get("MVF_SIDE_STREAM", "2") != "0" else None)
        self.side2 = (torch.cuda.Stream(device=device)
                      if self.side is not None and os.environ.get("MVF_SIDE_STREAM", "2") == "2" else None)
        if self.flat is not None and distributed and os.environ.get("MVF_OVERLAP_ALLREDUCE", "0") != "0":
            # MVF_OVERLAP_ALLREDUCE=<buckets>: the gradient arena is reduced bucket by bucket on a communication stream while the
            # backward still runs (optim.FlatAdamW.enable_overlap).  Off by default: see DESIGN.md section 5 for the measurement.
            self.flat.enable_overlap([self.stream, self.side, self.side2], n_buckets=max(2, int(os.environ["MVF_OVERLAP_ALLREDUCE"])))
        if self.side2 is not None and hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            # the shared pose weights receive gradients from two streams on purpose
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)

        # MVF_FILTER_BANK=1: every filter bank of the step packed by ONE launch at its start (conv_tc.FilterBank) instead of one small
        # launch per layer and direction.  Measured on B200 (config 2, same box, 30 steps): 9.78-9.83 ms with the bank, 9.65 ms
        # without -- the per-layer packs hide in the gaps of the concurrent streams while the single 50 us launch sits on the
        # critical path in front of the first convolution -- so per-layer packing stays the default.
        self.bank = None
        if device.type == "cuda" and os.environ.get("MVF_FILTER_BANK", "0") == "1":
            from . import conv, conv_tc
            if conv.get_backend() == "tcgen05":
                def conv_weights(mods):
                    seen, out = set(), []
                    for m in mods:
                        for sub in m.modules():
                            if isinstance(sub, (torch.nn.Conv2d, torch.nn.Linear)) and not isinstance(sub, torch.nn.ConvTranspose2d):
                                if getattr(sub, "groups", 1) == 1 and id(sub.weight) not in seen:
                                    seen.add(id(sub.weight))
                                    out.append(sub.weight)
                    return out
                self.bank = conv_tc.FilterBank(conv_weights(self.models.values()), conv_weights([self.vfi]) if self.vfi is not None else ())

    def train(self):
        for m in self.models.values():
            m.train()

    def set_lr(self, lr):
        """the scheduler's write to param_groups[..]["lr"] (train.py:289, 668); takes effect in the next step, also when
        that step is a CUDA-graph replay (the fused optimiser reads the rate from device memory)"""
        if self.flat is not None:
            self.flat.set_lr(lr)
        else:
            for g in self.optimizer.param_groups:
                g["lr"] = lr

    def forward_backward(self, inputs):
        if self.flat is not None:
            self.flat.zero_grad()
        elif self.reducer is not None:
            self.reducer.attach()  # zeroes the flat arena and points every .grad at its slice
        else:
            self.optimizer.zero_grad(set_to_none=True)
        if self.opt.multi_frame:
            out = multi_frame_losses(self.models, self.vfi, inputs, self.opt)
        else:
            out = single_frame_losses(self.models, inputs, self.opt, side=self.side, side2=self.side2)
        out["loss"].backward()
        # parameter gradients of the pose branch were produced on the side streams (single-frame step only: a stream that took
        # no part in the step must not be joined -- under graph capture that is a dependency on uncaptured work)
        for st in (() if self.opt.multi_frame else (self.side, self.side2)):
            if st is not None:
                torch.cuda.current_stream(self.device).wait_stream(st)
        if self.reducer is not None:
            self.reducer.allreduce_mean()
        return out

    def _step(self, inputs):
        if self.bank is not None:
            self.bank.refresh()
        out = self.forward_backward(inputs)
        if self.flat is not None:
            self.flat.step()  # all-reduce (N > 1) + clip + AdamW
        else:
            if self.opt.clip_grad is not None and self.opt.clip_grad > 0:
                torch.nn.utils.clip_grad_norm_([p for p in self.params + self.duplicated if p.grad is not None], self.opt.clip_grad)
            self.optimizer.step()
        loss = out["loss"].detach()
        for m in self.models.values():  # the reference's modules keep their last activations; drop the graph they hold
            if hasattr(m, "features"):
                m.features = None
            if hasattr(m, "outputs"):
                m.outputs = None
        return loss

    def __call__(self, inputs):
        if self.stream is None:
            return self._step(inputs)
        cur = torch.cuda.current_stream(self.device)
        if cur == self.stream:
            return self._step(inputs)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            loss = self._step(inputs)
        cur.wait_stream(self.stream)
        return loss


class GraphedTrainStep:
    """The whole optimisation step (zero_grad -> forward -> backward -> all-reduce -> clip -> AdamW) recorded once into
    a CUDA graph and replayed: one graph launch per step instead of several thousand kernel launches, so the GPU is
    never waiting for the host.  Inputs are copied into static device buffers before each replay (shapes are fixed, as
    they are in the reference's training loop: drop_last=True, train.py:110-117)."""

    def __init__(self, step, example_inputs, warmup=3):
        self.step = step
        dev = step.device
        self.static_inputs = {k: v.to(dev).clone() for k, v in example_inputs.items()}
        self.stream = step.stream
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):  # eager: lazy initialisations (workspaces, function attributes, used-parameter map)
                step(self.static_inputs)
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        if step.reducer is None and step.optimizer is not None:
            step.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.static_loss = step(self.static_inputs)

    def load(self, inputs):
        for k, v in inputs.items():
            self.static_inputs[k].copy_(v, non_blocking=True)

    def __call__(self, inputs=None):
        if inputs is not None:
            self.load(inputs)
        self.graph.replay()  # launched on the caller's current stream; the static buffers order it after load()
        # the replay changed every weight without touching the tensors' version counters: packed filter banks cached by an
        # eager forward outside the graph (validation, smoke tests) must not be reused afterwards
        from . import conv_tc
        conv_tc.weights_epoch += 1
        return self.static_loss


class HostFedRunner:
    """End-to-end driver for batches that live in pinned HOST memory: the H2D copy of batch i+1 runs on a copy stream
    while step i computes (two staging sets in HBM), the staged batch is moved into the graph's static inputs with one
    device-to-device copy, and the loss of every step is read back.  `feed(batch)` enqueues the next host batch,
    `run()` executes one step on the oldest staged batch and returns its loss as a Python float (D2H read)."""

    def __init__(self, graphed_step, example_host_batch):
        self.g = graphed_step
        dev = graphed_step.step.device
        self.dev = dev
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.stage = [{k: torch.empty_like(v, device=dev) for k, v in example_host_batch.items()} for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]   # H2D of the set finished
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]  # the set was copied into the static inputs
        self.head = self.tail = 0
        for e in self.consumed:
            e.record(torch.cuda.current_stream(dev))

    def feed(self, host_batch):
        i = self.head % 2
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[i])
            for k, v in host_batch.items():
                self.stage[i][k].copy_(v, non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self.head += 1

    def run(self):
        assert self.tail < self.head, "feed() a batch first"
        i = self.tail % 2
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.ready[i])
        self.g.load(self.stage[i])          # device-to-device, ~100 MB: tens of microseconds
        self.consumed[i].record(cur)
        self.tail += 1
        return self.g()                      # static loss tensor; float() it after feeding the next batch


def synthetic_inputs(opt, device=None, seed=1234, pin=False):
    """Synthetic 3-frame triplets of the named HxW: U[0,1) images, KITTI-normalised K, pinv(K)."""
    import numpy as np
    B, H, W = opt.batch_size, opt.height, opt.width
    g = torch.Generator().manual_seed(seed)
    inputs = {}
    for f in (-1, 0, 1):
        inputs[("color", f, 0)] = torch.rand(B, 3, H, W, generator=g)
        inputs[("color_aug", f, 0)] = torch.rand(B, 3, H, W, generator=g)
    K = np.array([[0.58 * W, 0, 0.5 * W, 0], [0, 1.92 * H, 0.5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    inv_K = np.linalg.pinv(K)
    inputs[("K", 0)] = torch.from_numpy(np.repeat(K[None], B, 0).copy())
    inputs[("inv_K", 0)] = torch.from_numpy(np.repeat(inv_K[None], B, 0).astype(np.float32).copy())
    if getattr(opt, "use_affine", False):
        # what datasets/mono_dataset.py:110-162 adds per item: a random rotation + scaled crop, the intrinsics-space
        # conjugation matrix Rc, the box in original-resolution pixels, and the two validity masks
        from . import affine as A
        angle = (torch.rand(B, 1, generator=g) * 2 - 1) * 10.0
        ratio = 1.2 + 0.8 * torch.rand(B, 1, generator=g)
        inputs["Rc"], inputs["box"] = A.augmentation_geometry(inputs[("K", 0)], inputs[("inv_K", 0)], angle, ratio,
                                                              torch.rand(B, 2, generator=g), H, W)
        inputs["ratio_local"], inputs["angle"] = ratio, angle
        ones = torch.ones(B, 1, H, W)
        rec = (A.affine_transform(ones, angle, inputs["box"]) > 0).float()
        inputs["valid_mask_rec"] = rec
        inputs["valid_mask_cons"] = (A.depth_restore(rec, angle, inputs["box"], torch.ones(B, 1)) > 0).float()
        for f in (-1, 0, 1):
            inputs[("color_affine", f, 0)] = A.affine_transform(inputs[("color", f, 0)], angle, inputs["box"])
        inputs[("color_affine_aug", 0, 0)] = A.affine_transform(inputs[("color_aug", 0, 0)], angle, inputs["box"])
    if pin:
        inputs = {k: v.pin_memory() for k, v in inputs.items()}
    if device is not None:
        inputs = {k: v.to(device) for k, v in inputs.items()}
    return inputs
