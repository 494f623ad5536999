#!/usr/bin/env python
"""Headline benchmark: training images/sec of the Mono-ViFI self-supervised inner loop (ResNet18 depth + pose,
192x640, batch 12 per GPU, synthetic 3-frame triplets), plus achieved HBM GB/s of the fused warp+SSIM kernel.

  python bench.py --gpus N --steps K --warmup W            our arm (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N ...            the reference path on the host CPU cores (rank 0 only)
  python bench.py --impl torch-eager ...                   the same step on stock eager PyTorch / cuDNN on the GPU (informational)
  python bench.py --config mf|3|4|5 ...                    the other BASELINE configurations (full multi-frame step; the default
                                                           N=1 run of config 2 also reports them under "extra_configs")

One JSON line on stdout (rank 0).  A "step" = zero_grad -> forward (2 pose nets, depth encoder + decoder, fused
view synthesis + photometric loss) -> backward -> gradient all-reduce (N>1) -> clip -> AdamW, i.e. the single-frame
slice of train.py:656-666 / 728-750 that BASELINE.json configs[1] names.
  value : device-timed, inputs resident in HBM          e2e : same step fed from pinned host memory each step
                                                              (H2D inside the timed region, loss read back)
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "training images/sec at 192x640 ResNet18, 1/2/4/8 B200; warp+SSIM HBM GB/s"
UNIT = "images/s"
WORKLOAD = "ResNet18 single-frame 192x640 batch 12 per GPU (BASELINE configs[1]): 2 pose nets + depth enc/dec + fused warp/SSIM loss group, fwd+bwd+AdamW"
MF = ("full multi-frame process_batch (train.py:698-885): 3 frozen IFRNet passes, 6 pose passes, single-frame + fused multi-frame depth of "
      "3 targets, 6 fused warp/SSIM loss groups, 3 SI-log consistency terms, fwd+bwd+AdamW")
# BASELINE.json configs[1..4] (+ the full multi-frame step of the headline model); Options overrides of trainer.Options
CONFIGS = {
    "2": dict(workload=WORKLOAD, opt=dict(batch_size=12, height=192, width=640)),
    "mf": dict(workload="ResNet18 192x640 batch 12 per GPU, " + MF + ", IFRNet-large as train.py:210 hard-codes",
               opt=dict(batch_size=12, height=192, width=640, multi_frame=True)),
    "3": dict(workload="D-HRNet (HRNet18) 192x640 batch 12 per GPU (BASELINE configs[2]), " + MF + ", IFRNet-large",
              opt=dict(batch_size=12, height=192, width=640, multi_frame=True, backbone="DHRNet")),
    "4": dict(workload="Lite-Mono 320x1024 batch 6 per GPU (configs/litemono/LiteMono_KITTI_HR.txt:13) + fusion_module + IFRNet_S "
                       "(BASELINE configs[3]), " + MF,
              opt=dict(batch_size=6, height=320, width=1024, multi_frame=True, backbone="LiteMono", vfi_scale="small")),
    "5": dict(workload="D-HRNet 384x1280 batch 8 per GPU (BASELINE configs[4]), " + MF + ", IFRNet-large",
              opt=dict(batch_size=8, height=384, width=1280, multi_frame=True, backbone="DHRNet")),
}
F1_FWD_BYTES_PER_PX = 40.0 + 8.0   # disp 4 + tgt 12 + 2 x src 12 (+ tie-break noise 8)  SURVEY.md 8(d)
F1_BWD_BYTES_PER_PX = 44.0         # re-read 40, write grad_disp 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-eager"])
    ap.add_argument("--config", default="2", choices=["2", "mf", "3", "4", "5"],
                    help="BASELINE.json configs: 2 = ResNet18 single-frame B12 192x640 (headline), mf = the full multi-frame ResNet18 step, "
                         "3 = D-HRNet full step B12 192x640, 4 = Lite-Mono 320x1024 + IFRNet_S B6, 5 = D-HRNet 384x1280 B8 multi-frame")
    ap.add_argument("--no-extra", action="store_true", help="skip the informational extra_configs / torch_eager_gpu legs of the N=1 run")
    ap.add_argument("--no-sync-bn", action="store_true", help="N>1: per-GPU BatchNorm statistics instead of SyncBatchNorm (A/B)")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's)")
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--cpu-batch", type=int, default=12, help="CPU arm: images per CPU step (the config's batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-fp32", action="store_true", help="e2e leg: feed the six fp32 tensors of the reference's loader instead of uint8 frames + GPU input pipeline")
    ap.add_argument("--torch-optimizer", action="store_true", help="torch clip_grad_norm_ + AdamW instead of the fused flat-arena kernels")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--conv-backend", default=os.environ.get("MVF_CONV_BACKEND", "tcgen05"), choices=["tcgen05", "cudnn"])
    args = ap.parse_args()
    o = CONFIGS[args.config]["opt"]
    args.batch = args.batch or o["batch_size"]
    args.height = args.height or o["height"]
    args.width = args.width or o["width"]
    args.workload = CONFIGS[args.config]["workload"]
    return args


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's path on host cores.  The reference is Python over torch and is not present on the GPU
# box, so this is the PORT: the same networks as plain torch modules on CPU (ATen convolutions, all host threads)
# and the C oracle (oracle/f1_oracle.c) for view synthesis + photometric loss, one thread per sample.
# ------------------------------------------------------------------------------------------------------------
def _cpu_step_factory(B, H, W):
    import numpy as np
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from oracle import f1_oracle as O
    from mono_vifi_b200 import layers as L, trainer as TR

    pool = ThreadPoolExecutor(max_workers=B)

    class OracleLoss(torch.autograd.Function):
        @staticmethod
        def forward(ctx, disp, P0, P1, tgt, s0, s1, inv_K, noise):
            a = [t.detach().numpy() for t in (disp, tgt, s0, s1, inv_K, P0, P1, noise)]

            def one(b):
                sl = [x[b:b + 1] for x in a]
                return O.f1_forward(sl[0], sl[1], sl[2], sl[3], sl[4], sl[5], sl[6], sl[7], None, full=False)
            outs = list(pool.map(one, range(B)))
            ctx.a, ctx.idx = a, [o["idx"] for o in outs]
            return torch.tensor(float(np.mean([o["loss"][0] for o in outs])), dtype=torch.float32)

        @staticmethod
        def backward(ctx, g):
            a, go = ctx.a, float(g) / B

            def one(b):
                sl = [x[b:b + 1] for x in a]
                return O.f1_backward(sl[0], sl[1], sl[2], sl[3], sl[4], sl[5], sl[6], ctx.idx[b], None, go)
            outs = list(pool.map(one, range(B)))
            gd = torch.from_numpy(np.concatenate([o[0] for o in outs], 0))
            gP0 = torch.from_numpy(np.concatenate([o[1] for o in outs], 0))
            gP1 = torch.from_numpy(np.concatenate([o[2] for o in outs], 0))
            return gd, gP0, gP1, None, None, None, None, None

    opt = TR.Options(batch_size=B, height=H, width=W)
    torch.manual_seed(1234)
    models = TR.build_models(opt, torch.device("cpu"))
    for m in models.values():
        m.train()
    params = [p for m in models.values() for p in m.parameters()]
    optim = torch.optim.AdamW(params, lr=opt.learning_rate, weight_decay=opt.weight_decay)
    inputs = TR.synthetic_inputs(opt)

    def step():
        optim.zero_grad(set_to_none=True)
        K, inv_K = inputs[("K", 0)], inputs[("inv_K", 0)]
        _, pose_0_n1 = TR.predict_poses(models, inputs[("color_aug", -1, 0)], inputs[("color_aug", 0, 0)])
        pose_0_p1, _ = TR.predict_poses(models, inputs[("color_aug", 0, 0)], inputs[("color_aug", 1, 0)])
        disp = models["depth"](models["encoder"](inputs[("color_aug", 0, 0)]))[("disp", 0)]
        P0, P1 = L.matmul_KT(K, pose_0_n1)[:, :3], L.matmul_KT(K, pose_0_p1)[:, :3]
        noise = torch.randn(B, 2, H, W)
        loss = OracleLoss.apply(disp, P0, P1, inputs[("color", 0, 0)], inputs[("color", -1, 0)],
                                inputs[("color", 1, 0)], inv_K, noise)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], opt.clip_grad)
        optim.step()
        return float(loss)
    return step


def time_cpu(B, H, W, steps, warmup, budget_s=25.0):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = _cpu_step_factory(B, H, W)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    n = 0
    while n < steps:
        step()
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": B * n / dt, "ms_per_step": 1e3 * dt / n, "steps": n, "cores": cores,
            "sample": "%d steps of batch %d at %dx%d (same step, bounded batch)" % (n, B, H, W)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = time_cpu(args.cpu_batch, args.height, args.width, args.steps, max(1, min(args.warmup, 2)), budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": max(1, min(args.warmup, 2)), "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIGS["2"]["workload"], "height": args.height, "width": args.width,
                       "note": "reference path on host CPU cores: ATen convolutions (all threads) + C oracle of the "
                               "view-synthesis/photometric loss, bounded sample of batch %d per step" % args.cpu_batch},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for nme, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mono_vifi_b200 import _lib, bn_act, conv, conv_tc, ddp, fused, trainer as TR

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: there is no CPU fallback")
    _lib.lib()  # fail loudly if the CUDA extension is missing
    rank, local, world = ddp.init_from_env("nccl")
    if world != args.gpus and rank == 0:
        sys.stderr.write("warning: --gpus %d but WORLD_SIZE %d\n" % (args.gpus, world))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    conv.set_backend(args.conv_backend)
    okw = dict(CONFIGS[args.config]["opt"], batch_size=args.batch, height=args.height, width=args.width)
    if args.no_sync_bn:
        okw["sync_bn"] = False
    opt = TR.Options(**okw)
    torch.manual_seed(1234)
    step = TR.TrainStep(opt, dev, distributed=(world > 1), capturable=not args.no_graph,
                        fused_optimizer=not args.torch_optimizer)
    step.train()
    ddp.broadcast_parameters(step.params)
    # two distinct synthetic batches per rank, rotated, in pinned host memory and (for `value`) resident in HBM
    host = [TR.synthetic_inputs(opt, seed=1234 + 17 * rank + s, pin=True) for s in range(2)]
    resident = [{k: v.to(dev) for k, v in h.items()} for h in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- eager steps: lazy initialisation, launch accounting and per-kernel CUDA-event timing -------------------
    for i in range(max(1, args.warmup - 2)):
        step(resident[i % 2])
    fused.timing, conv_tc.timing = [], []
    l0, c0, b0 = dict(fused.launches), dict(conv_tc.launches), dict(bn_act.launches)
    for k in conv.stats:
        conv.stats[k] = 0
    n_eager = 3
    # per-kernel event timing wants the kernels alone on the GPU: no concurrent pose branches
    (side, side2), step.side, step.side2 = (step.side, step.side2), None, None
    marks = [0]
    for i in range(n_eager):
        # the GPU first spins for ~0.25 s (eager host time of the largest configuration's step is below that) so that the host
        # enqueues the whole step behind it: every event interval below is then GPU time of one kernel, not launch latency
        torch.cuda._sleep(int(5e8))
        step(resident[i % 2])
        marks.append(len(conv_tc.timing))
    torch.cuda.synchronize()
    step.side, step.side2 = side, side2
    kt = {"f1_fwd": [], "f1_bwd": []}
    for tag, a, b in fused.timing:
        kt[tag].append(a.elapsed_time(b))
    ct = {}
    per_step = [conv_tc.timing[marks[i]:marks[i + 1]] for i in range(n_eager)]
    same_seq = all(len(p) == len(per_step[0]) and [(t[0], t[1]) for t in p] == [(t[0], t[1]) for t in per_step[0]] for p in per_step)
    if same_seq:
        # every step launches the same sequence: a launch's time is its fastest of the n_eager steps.  (An interval is GPU time only
        # while the host stays ahead of the GPU; a host hiccup between the two event records -- an allocator miss, a lazy module
        # load -- lands inside one interval of one step: once 5 ms on a 14 us kernel, profiles/README.md.)
        for j, (tag, fl, _, _) in enumerate(per_step[0]):
            ent = ct.setdefault(tag, [0.0, 0.0, 0])
            ent[0] += fl * n_eager
            ent[1] += min(p[j][2].elapsed_time(p[j][3]) for p in per_step) * 1e-3 * n_eager
            ent[2] += n_eager
    else:
        for tag, fl, a, b in conv_tc.timing:
            ent = ct.setdefault(tag, [0.0, 0.0, 0])
            ent[0] += fl
            ent[1] += a.elapsed_time(b) * 1e-3
            ent[2] += 1
    fused.timing = conv_tc.timing = None
    conv_launches = {k: (conv_tc.launches[k] - c0[k]) // n_eager for k in c0}
    conv_calls = {k: v // n_eager for k, v in conv.stats.items()}
    bn_calls = {k: (bn_act.launches[k] - b0[k]) // n_eager for k in b0}
    # F1 + convolution kernels + 3 kernels per fused BN call + (sum of squares, AdamW, tick, gradient gather) of the optimiser;
    # the decoder's upcat / max-pool / activation-backward kernels are not counted (under-claim)
    launches_per_step = (sum(fused.launches[k] - l0[k] for k in l0) // n_eager + sum(conv_launches.values()) + 3 * sum(bn_calls.values()) +
                         (0 if args.torch_optimizer else 5))
    # ---- the step as a CUDA graph (one launch per step) ---------------------------------------------------------
    graph_note = "eager launches (--no-graph)"
    run = step
    if not args.no_graph:
        try:
            run = TR.GraphedTrainStep(step, resident[0], warmup=2)
            graph_note = "whole step recorded once into a CUDA graph and replayed"
        except Exception as e:  # keep measuring: eager is the same arithmetic
            run = step
            sys.stderr.write("CUDA graph capture failed: %s\n" % e)
            graph_note = "CUDA graph capture failed (%s): eager launches" % str(e).splitlines()[0][:120]
    for i in range(2):
        run(resident[i % 2])
    # ---- timed region 1: device-resident inputs ---------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = run(resident[i % 2])
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1) / args.steps
    my_launches = launches_per_step * args.steps
    # ---- timed region 2: end to end from pinned host memory -----------------------------------------------------
    # The host hands over what a loader produces after its resize: the item's three uint8 frames + the drawn augmentation
    # parameters (+ K, inv_K); ToTensor / flip / ColorJitter run on the GPU (mono_vifi_b200/input_pipeline.py, mvf_input_pipeline)
    # and write straight into the step's inputs.  --e2e-fp32 feeds the six fp32 tensors of the reference's loader instead.
    import numpy as np
    from mono_vifi_b200 import input_pipeline as IP
    if args.e2e_fp32:
        host_e2e = host
    else:
        host_e2e = []
        for sidx in range(2):
            rng = np.random.RandomState(1234 + 17 * rank + sidx)
            pf, pi = IP.draw_params(args.batch, rng)
            host_e2e.append({"frames_u8": torch.from_numpy(rng.randint(0, 256, (args.batch, 3, args.height, args.width, 3)).astype(np.uint8)).pin_memory(),
                             "jitter_f": pf.pin_memory(), "jitter_i": pi.pin_memory(),
                             ("K", 0): host[sidx][("K", 0)], ("inv_K", 0): host[sidx][("inv_K", 0)]})
    h2d_bytes = sum(v.numel() * v.element_size() for v in host_e2e[0].values())
    stage = [{k: torch.empty_like(v, device=dev) for k, v in host_e2e[0].items()} for _ in range(2)]
    eager_buf = {k: torch.empty_like(v, device=dev) for k, v in host[0].items()}
    pipe = None if args.e2e_fp32 else IP.InputPipeline(args.batch, args.height, args.width, dev)
    feeder = None
    if run is not step:
        feeder = TR.HostFedRunner(run, host[0]) if args.e2e_fp32 else IP.U8HostFedRunner(run, host_e2e[0])
        feeder.feed(host_e2e[0])
        feeder.run()                            # untimed: lazy initialisation of the feeding path
        torch.cuda.synchronize()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = 0.0
    if feeder is not None:
        feeder.feed(host_e2e[0])                # inside the timed region: every step's H2D is
    for i in range(args.steps):
        if run is step:
            buf = stage[i % 2]
            for k, v in host_e2e[i % 2].items():
                buf[k].copy_(v, non_blocking=True)
            if pipe is not None:
                pipe(buf["frames_u8"], buf["jitter_f"], buf["jitter_i"], out=eager_buf)
                eager_buf[("K", 0)].copy_(buf[("K", 0)])
                eager_buf[("inv_K", 0)].copy_(buf[("inv_K", 0)])
                buf = eager_buf
            last = float(step(buf))  # D2H read of the step's loss
        else:
            loss_t = feeder.run()               # step i on the staged batch (graph replay)
            if i + 1 < args.steps:
                feeder.feed(host_e2e[(i + 1) % 2])  # H2D of batch i+1 on the copy stream, overlapping step i
            last = float(loss_t)                # D2H read of the step's loss (synchronises)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3) / args.steps
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return 0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    px = args.batch * args.height * args.width

    traffic = {}
    try:  # DRAM bytes per launch from the committed `ncu --set full` capture (only valid for the benchmark shape)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
    except Exception:
        pass
    std_shape = (args.batch, args.height, args.width) == (12, 192, 640)

    def roof(tag, bpp):
        v = sorted(kt[tag])
        if not v:
            return None
        avg_ms = sum(v) / len(v)
        ach = bpp * px / (avg_ms * 1e-3) / 1e9
        tr = traffic.get(tag + "_kernel", {}).get("dram_bytes_per_launch") if std_shape else None
        return {"kernel": (traffic.get(tag + "_kernel") or {}).get("kernel", tag + "_kernel") if std_shape else tag + "_kernel", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                "traffic": tr, "avg_launch_us": avg_ms * 1e3, "launches_timed": len(v), "bytes_per_px": bpp,
                "algorithmic_bytes_per_launch": bpp * px, "peak_source": peak_src,
                "note": "instruction-issue / latency bound (about 23 useful warp-instructions per pixel), see DESIGN.md section 4 and "
                        "profiles/r2_f1_ablation.md / r2_f1_variants.md; traffic = dram__bytes of the committed ncu capture of this kernel and shape"}
    tf32_peak = float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2.0
    conv_roof = {}
    for tag, (fl, sec, n) in sorted(ct.items()):
        ach = fl / sec / 1e12 if sec > 0 else 0.0
        conv_roof[tag] = {"kernel": "conv_%s (tcgen05, tf32)" % tag, "bound": "tensor", "achieved": ach, "peak": tf32_peak,
                          "unit": "TFLOP/s", "frac": ach / tf32_peak, "launches_timed": n, "launches_per_step": n // n_eager, "flop_per_step": fl / n_eager,
                          "ms_per_step": 1e3 * sec / n_eager,
                          "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (tf32 dense rate is half of bf16)"}
    line = {"metric": METRIC, "value": args.batch * world / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (tf32 tensor-core convolutions, as torch's cuDNN default)", "data": "synthetic",
            "config": {"workload": args.workload, "config_id": args.config, "per_gpu_batch": args.batch, "global_batch": args.batch * world,
                       "height": args.height, "width": args.width, "parallelism": "dp%d" % world,
                       "launch": graph_note, "streams": ("pose passes on side streams concurrent with the depth branch, weight gradients on companion streams"
                                   if step.side is not None else "one stream"),
                       "optimizer": "torch clip_grad_norm_ + AdamW" if args.torch_optimizer else "fused clip + AdamW over flat arenas (mvf_adamw_step)",
                       "conv_backend": conv.get_backend(), "conv_calls_per_step": conv_calls,
                       "conv_kernel_launches_per_step": conv_launches, "bn_calls_per_step": bn_calls,
                       "l2": "working set (activations, several GB) is far larger than the 126 MB L2; inputs rotate between two batches"},
            "e2e": {"value": args.batch * world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "feed": ("six fp32 tensors per item (the reference loader's output)" if args.e2e_fp32 else
                             "uint8 frames + augmentation parameters; ToTensor / flip / ColorJitter on the GPU (mvf_input_pipeline)")},
            "gpu_launches": my_launches, "clocks": clocks, "loss": last,
            "roofline": roof("f1_fwd", F1_FWD_BYTES_PER_PX), "roofline_bwd": roof("f1_bwd", F1_BWD_BYTES_PER_PX),
            "roofline_conv": conv_roof}
    if world == 1 and not args.no_cpu_baseline and args.config == "2":
        r = time_cpu(args.cpu_batch, args.height, args.width, 3, 1, budget_s=25.0)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    if world == 1 and not args.no_extra and args.config == "2":
        # informational legs (never the headline): the same step on stock eager PyTorch / cuDNN on this GPU, and the other
        # BASELINE configs, each in its own process so that a failure there cannot touch the line above
        del run, feeder, step
        torch.cuda.empty_cache()
        try:
            line["torch_eager_gpu"] = time_torch_eager(args.batch, args.height, args.width, dev, resident, 10, 3)
        except Exception as e:
            line["torch_eager_gpu"] = {"error": str(e).splitlines()[0][:200]}
        line["extra_configs"] = {c: run_extra_config(c) for c in ("mf", "3", "4", "5")}
    emit(line)
    return 0


def time_torch_eager(B, H, W, dev, batches, steps, warmup):
    """The single-frame step of configs[1] on stock torch (baseline/torch_eager.py): cuDNN convolutions with torch's default
    TF32 setting, nn.BatchNorm2d, F.grid_sample, unfused loss, torch AdamW, eager launches."""
    import torch
    from baseline import torch_eager as TE
    torch.manual_seed(1234)
    out = {}
    for tag, cl in (("nchw", False), ("channels_last", True)):
        st = TE.SingleFrameStep(B, H, W, dev, channels_last=cl)
        for i in range(warmup):
            st(batches[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            loss = st(batches[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[tag] = {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup, "loss": float(loss)}
        del st
        torch.cuda.empty_cache()
    best = max(out.values(), key=lambda r: r["value"])
    return {"value": best["value"], "unit": UNIT, "ms_per_step": best["ms_per_step"], "layouts": out,
            "what": "stock torch %s / cuDNN eager, same step, same batch, inputs resident (baseline/torch_eager.py)" % torch.__version__}


def run_extra_config(cfg, steps=5, warmup=3, timeout=170):
    cmd = [sys.executable, os.path.abspath(__file__), "--config", cfg, "--no-extra", "--no-cpu-baseline", "--steps", str(steps),
           "--warmup", str(warmup)]
    t0 = time.perf_counter()
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        rows = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
        if p.returncode != 0 or not rows:
            return {"error": "exit %d: %s" % (p.returncode, (p.stderr.strip().splitlines() or ["no output"])[-1][:300])}
        r = json.loads(rows[-1])
        keep = {k: r.get(k) for k in ("value", "unit", "ms_per_step", "steps", "warmup", "gpu_launches", "loss")}
        keep["e2e"] = (r.get("e2e") or {}).get("value")
        keep["workload"] = r["config"]["workload"]
        for k in ("launch", "conv_calls_per_step", "conv_kernel_launches_per_step", "bn_calls_per_step", "per_gpu_batch", "height", "width"):
            keep[k] = r["config"].get(k)
        keep["f1_fwd"] = {k: (r.get("roofline") or {}).get(k) for k in ("avg_launch_us", "frac", "launches_timed")}
        keep["f1_bwd"] = {k: (r.get("roofline_bwd") or {}).get(k) for k in ("avg_launch_us", "frac", "launches_timed")}
        keep["conv_frac"] = {k: v.get("frac") for k, v in (r.get("roofline_conv") or {}).items()}
        keep["wall_s"] = time.perf_counter() - t0
        return keep
    except subprocess.TimeoutExpired:
        return {"error": "timeout after %d s" % timeout}
    except Exception as e:
        return {"error": str(e)[:300]}


def run_torch_eager(args):
    """`--impl torch-eager`: the GPU eager-PyTorch arm alone (N=1), printed in the same line format."""
    import torch
    from mono_vifi_b200 import trainer as TR
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    dev = torch.device("cuda", 0)
    opt = TR.Options(batch_size=args.batch, height=args.height, width=args.width)
    batches = [TR.synthetic_inputs(opt, dev, seed=1234 + s) for s in range(2)]
    r = time_torch_eager(args.batch, args.height, args.width, dev, batches, args.steps, args.warmup)
    emit({"impl": "torch-eager", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f32 (cuDNN TF32 convolutions, torch default)", "data": "synthetic",
          "config": {"workload": CONFIGS["2"]["workload"], "what": r["what"], "layouts": r["layouts"]}, "gpu_launches": 0})
    return 0


_RESULT_FD = None


def _reserve_stdout():
    """stdout carries exactly one JSON line.  Native libraries (NCCL's version banner, cuDNN warnings) printf to fd 1,
    so fd 1 is pointed at stderr for the whole run and the result line is written to a private duplicate of the
    original stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    args = parse()
    _reserve_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "torch-eager":
        return run_torch_eager(args)
    rc = run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
    return rc


if __name__ == "__main__":
    sys.exit(main())
