"""Worker of tests/test_ddp_nccl.py (one process per GPU, launched by torch.distributed.run): the data-parallel step on a
batch sharded over the ranks must produce what ONE process produces on the whole batch --
  * every parameter gradient after the all-reduce (mean) == the single-process gradient on the global batch,
  * every BatchNorm buffer (running_mean / running_var / num_batches_tracked) identical on all ranks and == single process,
which is what the reference gets from SyncBatchNorm + DistributedDataParallel (train.py:205-208).
Prints one JSON line from rank 0."""
import copy
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mono_vifi_b200 import conv_tc, ddp, peer, trainer as TR  # noqa: E402


def main():
    rank, local, world = ddp.init_from_env("nccl")
    dev = torch.device("cuda", local)
    multi = len(sys.argv) > 1 and sys.argv[1] == "mf"
    per_rank = 2
    B = per_rank * world
    H, W = 64, 96
    conv_tc.precision.set("3xtf32")     # fp32-class arithmetic so that the comparison is tight
    opt_g = TR.Options(batch_size=B, height=H, width=W, tie_break_noise=False, multi_frame=multi, vfi_scale="small")
    opt_l = TR.Options(batch_size=per_rank, height=H, width=W, tie_break_noise=False, multi_frame=multi, vfi_scale="small")
    torch.manual_seed(21)
    base = TR.build_models(opt_g, dev)
    with torch.no_grad():   # pose outputs are ~1e-3 at initialisation: make the warp a real one
        base["pose"].convs[("pose", 2)].bias.normal_(0, 2.0)
    ddp.broadcast_module_state(base.values())
    inputs_g = TR.synthetic_inputs(opt_g, dev, seed=33)
    inputs_l = {k: v[rank::world].contiguous() for k, v in inputs_g.items()}   # CustomDistributedSampler's strided split
    torch.manual_seed(5)
    vfi_state = None
    # ---- single process, whole batch (every rank computes it; rank 0 reports) ----
    m_s = copy.deepcopy(base)
    step_s = TR.TrainStep(opt_g, dev, models=m_s, distributed=False)
    if multi:
        vfi_state = copy.deepcopy(step_s.vfi.state_dict())
        for t in vfi_state.values():
            dist.broadcast(t, 0)
        step_s.vfi.load_state_dict(vfi_state)
    step_s.train()
    step_s.side = step_s.side2 = None
    with torch.cuda.stream(step_s.stream):
        out_s = step_s.forward_backward(inputs_g)
    torch.cuda.synchronize()
    # ---- data parallel: this rank's shard, SyncBatchNorm, gradients averaged over the ranks ----
    m_d = copy.deepcopy(base)
    step_d = TR.TrainStep(opt_l, dev, models=m_d, distributed=True)
    if multi:
        step_d.vfi.load_state_dict(vfi_state)
    step_d.train()
    n_sync = sum(isinstance(mod, torch.nn.SyncBatchNorm) for m in m_d.values() for mod in m.modules())
    with torch.cuda.stream(step_d.stream):
        out_d = step_d.forward_backward(inputs_l)
    torch.cuda.synchronize()
    worst, worst_name, n = 0.0, None, 0
    worst_net, worst_net_name = 0.0, None
    num = den = 0.0
    per_param = []
    per_param_l2 = []
    for (name_s, mod_s), (name_d, mod_d) in zip(m_s.items(), m_d.items()):
        # the largest gradient entry of this network: the yardstick for parameters whose own gradient is orders of magnitude smaller
        # (their entries are sums that cancel to rounding level, e.g. the last ResNet block's 512x512x3x3 filters on a 2x3 map)
        net_scale = max([float(p.grad.abs().max()) for p in mod_s.parameters() if p.grad is not None] + [0.0])
        for (pn, ps), (_, pd) in zip(mod_s.named_parameters(), mod_d.named_parameters()):
            if ps.grad is None:
                assert pd.grad is None, (name_s, pn)
                continue
            g = pd.grad.detach().clone()
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            g /= world
            scale = float(ps.grad.abs().max())
            if scale > 0:
                e = float((g - ps.grad).abs().max()) / scale
                per_param.append((e, "%s.%s" % (name_s, pn)))
                if os.environ.get("MVF_DDP_DEBUG") and pn.endswith("layer4.1.conv2.weight") and rank == 0:
                    print("DEBUG %s.%s single: norm %.9g first %s | dp: norm %.9g first %s" % (
                        name_s, pn, float(ps.grad.double().norm()), [round(float(v), 9) for v in ps.grad.flatten()[:3]],
                        float(g.double().norm()), [round(float(v), 9) for v in g.flatten()[:3]]), flush=True)
                if e > worst:
                    worst, worst_name = e, "%s.%s" % (name_s, pn)
                e_net = float((g - ps.grad).abs().max()) / max(scale, 1e-3 * net_scale)
                if e_net > worst_net:
                    worst_net, worst_net_name = e_net, "%s.%s" % (name_s, pn)
                l2 = float((g - ps.grad).double().norm()) / max(float(ps.grad.double().norm()), 1e-30)
                per_param_l2.append((l2, "%s.%s" % (name_s, pn), scale, net_scale))
            num += float((g - ps.grad).double().pow(2).sum())
            den += float(ps.grad.double().pow(2).sum())
            n += 1
    buf_err, buf_spread = 0.0, 0.0
    for (name_s, mod_s), (_, mod_d) in zip(m_s.items(), m_d.items()):
        for (bn, bs), (_, bd) in zip(mod_s.named_buffers(), mod_d.named_buffers()):
            a, b = bs.double(), bd.double()
            buf_err = max(buf_err, float((a - b).abs().max()) / max(1e-6, float(a.abs().max())))
            lo, hi = b.clone(), b.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            buf_spread = max(buf_spread, float((hi - lo).abs().max()))
    loss_d = out_d["loss"].detach().clone()
    dist.all_reduce(loss_d, op=dist.ReduceOp.SUM)
    loss_d /= world
    # ---- one full optimiser step through the flat arena (all-reduce inside FlatAdamW.step): weights stay identical ----
    with torch.cuda.stream(step_d.stream):
        step_d.flat.step()
    torch.cuda.synchronize()
    # ---- a second, complete step on both sides: the data-parallel one now reduces its gradient buckets on the communication
    # stream during the backward (FlatAdamW.enable_overlap takes effect once the arena exists) ----
    with torch.cuda.stream(step_s.stream):
        step_s.flat.step()
        step_s._step(inputs_g)
    with torch.cuda.stream(step_d.stream):
        step_d._step(inputs_l)
    torch.cuda.synchronize()
    ov = step_d.flat._overlap
    buckets_overlapped = 0 if ov is None else len(ov["bounds"])
    # the reduced gradient arena of the second step (sum over ranks; the 1 / world scale lives in the AdamW kernel) against the single
    # process's arena.  (Weights themselves are a poor yardstick: Adam's first steps move every weight by ~lr whatever the size of its
    # gradient, so parameters with noise-level gradients differ by whole updates between any two runs.)
    gs, gd = step_s.flat.G.double(), step_d.flat.G.double() / world
    w_err_rel = float((gs - gd).norm() / gs.norm()) if gs.numel() == gd.numel() else float("nan")
    w_spread = 0.0
    for m in m_d.values():
        for p in m.parameters():
            lo, hi = p.detach().clone(), p.detach().clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            w_spread = max(w_spread, float((hi - lo).abs().max()))
    if rank == 0:
        print(json.dumps({"world": world, "multi_frame": multi, "sync_bn_modules": n_sync, "params_compared": n,
                          "grad_worst_rel": worst, "grad_worst_name": worst_name,
                          "grad_worst_rel_floored": worst_net, "grad_worst_floored_name": worst_net_name,
                          "grad_worst_param_rel_l2": max(per_param_l2)[0] if per_param_l2 else 0.0,
                          "grad_worst_five_rel_l2": [[float("%.3g" % e), nm, float("%.3g" % sc), float("%.3g" % ns)] for e, nm, sc, ns in sorted(per_param_l2, reverse=True)[:5]],
                          "grad_worst_five": [[round(e, 6), nm] for e, nm in sorted(per_param, reverse=True)[:5]], "grad_rel_l2": (num / max(den, 1e-30)) ** 0.5,
                          "buffer_rel_err_vs_single": buf_err, "buffer_spread_across_ranks": buf_spread,
                          "loss_single": float(out_s["loss"]), "loss_dp_mean": float(loss_d),
                          "weight_spread_after_step": w_spread, "peer_exchanges": peer.exchanges,
                          "allreduce_buckets_overlapped": buckets_overlapped, "second_step_reduced_grad_rel_l2": w_err_rel,
                          "exchange": "nvlink peer memory" if peer.exchanges else "nccl"}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
