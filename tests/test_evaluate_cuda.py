"""GPU: the evaluation path (csrc/eval.cu through mvf_depth_eval / mvf_bn_eval_fwd, mono_vifi_b200/evaluate.py) against the
reference's own loop (train.py:436-468: F.interpolate -> 1 / disp -> mask / Eigen crop -> torch.median scaling -> clamp ->
compute_depth_errors, layers.py:293-311) written with the same torch calls, on synthetic LiDAR-like ground truth."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _reference_image(pred_disp, gt_depth, eigen, stereo):
    import torch
    import torch.nn.functional as F
    from mono_vifi_b200 import layers as L
    gh, gw = gt_depth.shape
    pd = F.interpolate(pred_disp[None, None], (gh, gw), mode="bilinear", align_corners=False)
    pred_depth = 1 / pd[0, 0]
    if eigen:
        mask = (gt_depth > 1e-3) & (gt_depth < 80)
        crop = torch.zeros_like(mask)
        crop[int(0.40810811 * gh):int(0.99189189 * gh), int(0.03594771 * gw):int(0.96405229 * gw)] = 1
        mask = mask * crop
    else:
        mask = gt_depth > 0
    p, g = pred_depth[mask], gt_depth[mask]
    ratio = torch.tensor(5.4, device=p.device) if stereo else torch.median(g) / torch.median(p)
    p = torch.clamp(p * ratio, 1e-3, 80)
    return torch.stack(list(L.compute_depth_errors(g.double(), p.double())) + [ratio.double()])


@pytest.mark.parametrize("gh,gw,h,w,eigen,stereo", [(375, 1242, 192, 640, True, False), (370, 1226, 320, 1024, True, False),
                                                    (375, 1242, 192, 640, False, False), (376, 1241, 192, 640, True, True),
                                                    (64, 96, 32, 48, True, False)])
def test_depth_metrics_match_the_reference_loop(gh, gw, h, w, eigen, stereo):
    import torch
    from mono_vifi_b200 import evaluate as E
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(gh + w)
    pred_disp = 0.01 + 0.3 * torch.rand(h, w, device=dev, generator=g)
    gt = 1.0 + 90.0 * torch.rand(gh, gw, device=dev, generator=g)
    gt = torch.where(torch.rand(gh, gw, device=dev, generator=g) < 0.05, gt, torch.zeros_like(gt))   # ~5 % valid LiDAR returns
    out = E.depth_metrics(pred_disp, gt, "eigen" if eigen else "benchmark", stereo).double()
    ref = _reference_image(pred_disp, gt, eigen, stereo)
    # the median is exact (radix select); sums differ from torch's only by fp32 vs fp64 accumulation of the per-pixel terms
    assert float((out[7] - ref[7]).abs()) <= 2e-6 * float(ref[7].abs())
    assert float(((out[:7] - ref[:7]).abs() / ref[:7].abs().clamp_min(1e-6)).max()) <= 1e-4, (out, ref)
    assert torch.equal(out, E.depth_metrics(pred_disp, gt, "eigen" if eigen else "benchmark", stereo).double())


def test_evaluate_depth_runs_the_networks_in_inference_mode():
    import torch
    from mono_vifi_b200 import bn_act, conv, evaluate as E, trainer as TR
    dev = torch.device("cuda:0")
    opt = TR.Options(batch_size=2, height=64, width=96)
    torch.manual_seed(4)
    models = TR.build_models(opt, dev)
    for m in models.values():   # a few training steps' worth of running statistics
        m.train()
    g = torch.Generator().manual_seed(8)
    imgs = [torch.rand(2, 3, 64, 96, generator=g) for _ in range(3)]
    with torch.no_grad():
        for x in imgs:
            models["depth"](models["encoder"](x.to(dev)))
    gts = [np.where(np.random.RandomState(i).rand(120, 200) < 0.1, 1 + 70 * np.random.RandomState(i + 9).rand(120, 200), 0).astype(np.float32)
           for i in range(6)]
    n0, c0 = bn_act.launches.get("bn_eval", 0), dict(conv.stats)
    res = E.evaluate_depth({"encoder": models["encoder"], "depth": models["depth"]}, [{("color", 0, 0): x} for x in imgs], gts, opt)
    assert res["n_images"] == 6 and all(np.isfinite(res[k]) for k in E.METRICS) and 0 <= res["a1"] <= res["a2"] <= res["a3"] <= 1
    assert bn_act.launches["bn_eval"] - n0 == 3 * 20 and conv.stats["cudnn"] == c0["cudnn"]     # 20 BatchNorms of ResNet18, all on the kernel
    assert models["encoder"].training                                                           # restored (train.py:482 set_train)
    # the same numbers from the reference's loop on the same predictions
    for m in models.values():
        m.eval()
    import torch.nn.functional as F  # noqa: F401
    from mono_vifi_b200 import layers as L
    refs = []
    with torch.no_grad():
        k = 0
        for x in imgs:
            pd, _ = L.disp_to_depth(models["depth"](models["encoder"](x.to(dev)))[("disp", 0)], opt.min_depth, opt.max_depth)
            for b in range(2):
                refs.append(_reference_image(pd[b, 0], torch.from_numpy(gts[k]).to(dev), True, False))
                k += 1
    ref = torch.stack(refs).mean(0)
    for j, name in enumerate(E.METRICS):
        assert abs(res[name] - float(ref[j])) <= 1e-4 * max(abs(float(ref[j])), 1e-6), (name, res[name], float(ref[j]))
