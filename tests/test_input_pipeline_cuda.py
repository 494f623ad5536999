"""GPU: the input-pipeline kernel (csrc/input.cu through mvf_input_pipeline) against torchvision's own tensor implementation of
ToTensor / hflip / ColorJitter (transforms.functional.adjust_* applied in the drawn order) on the same uint8 frames: fp32 agreement
(1e-4: FMA contraction and the order of the grey-mean reduction), and against the PIL path the reference's loader runs.  torchvision's
OWN two backends differ there: 8-bit rounding after every operation (~1/255 each for brightness / contrast / saturation) and PIL's
integer HSV conversion (up to 9.6/255 for adjust_hue alone, measured tensor-vs-PIL on random images) -- the chain of four measured 16/255 here; the bound (20/255) only guards against gross errors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tv_chain(img, order, f):
    import torchvision.transforms.functional as TF
    for op in order:
        if op == 0:
            img = TF.adjust_brightness(img, f[0])
        elif op == 1:
            img = TF.adjust_contrast(img, f[1])
        elif op == 2:
            img = TF.adjust_saturation(img, f[2])
        else:
            img = TF.adjust_hue(img, f[3])
    return img


def test_input_pipeline_matches_torchvision():
    import torch
    from PIL import Image
    from mono_vifi_b200 import input_pipeline as IP
    dev = torch.device("cuda:0")
    B, H, W = 5, 48, 80
    rng = np.random.RandomState(3)
    # smooth + noisy content so that hue / saturation see real colour
    yy, xx = np.mgrid[0:H, 0:W]
    frames = np.stack([np.stack([np.stack([(127 + 100 * np.sin(0.1 * xx * (c + 1) + b + f) * np.cos(0.07 * yy + c)) for c in range(3)], -1)
                                 for f in range(3)]) for b in range(B)])
    frames = np.clip(frames + rng.randint(-20, 20, frames.shape), 0, 255).astype(np.uint8)        # [B,3,H,W,3]
    prm_f, prm_i = IP.draw_params(B, rng)
    prm_i[0, 4], prm_i[0, 5] = 1, 1
    prm_i[1, 4], prm_i[1, 5] = 1, 0
    prm_i[2, 4], prm_i[2, 5] = 0, 1
    prm_i[3, 4] = 1
    prm_i[3, :4] = torch.tensor([3, 1, 0, 2])            # contrast in the middle of the chain
    pipe = IP.InputPipeline(B, H, W, dev)
    out = pipe(torch.from_numpy(frames).to(dev), prm_f.to(dev), prm_i.to(dev))
    worst_tv = worst_pil = 0.0
    for b in range(B):
        order, f = [int(v) for v in prm_i[b, :4]], [float(v) for v in prm_f[b]]
        for k, fid in enumerate((-1, 0, 1)):
            img = torch.from_numpy(frames[b, k]).permute(2, 0, 1).float() / 255.0
            pil = Image.fromarray(frames[b, k])
            if prm_i[b, 5]:
                img, pil = img.flip(-1), pil.transpose(Image.FLIP_LEFT_RIGHT)
            assert torch.equal(out[("color", fid, 0)][b].cpu(), img)
            aug, aug_pil = img, pil
            if prm_i[b, 4]:
                aug, aug_pil = _tv_chain(img, order, f), _tv_chain(pil, order, f)
            worst_tv = max(worst_tv, float((out[("color_aug", fid, 0)][b].cpu() - aug).abs().max()))
            ref_pil = torch.from_numpy(np.asarray(aug_pil)).permute(2, 0, 1).float() / 255.0
            worst_pil = max(worst_pil, float((out[("color_aug", fid, 0)][b].cpu() - ref_pil).abs().max()))
    assert worst_tv <= 1e-4, worst_tv
    assert worst_pil <= 20.0 / 255.0, worst_pil


def test_u8_host_fed_runner_feeds_a_captured_step():
    import torch
    from mono_vifi_b200 import input_pipeline as IP, trainer as TR
    dev = torch.device("cuda:0")
    opt = TR.Options(batch_size=2, height=64, width=96)
    torch.manual_seed(0)
    step = TR.TrainStep(opt, dev)
    step.train()
    g = TR.GraphedTrainStep(step, TR.synthetic_inputs(opt, dev), warmup=2)
    rng = np.random.RandomState(0)
    host = []
    for _ in range(2):
        f, i = IP.draw_params(2, rng)
        ex = TR.synthetic_inputs(opt)
        host.append({"frames_u8": torch.from_numpy(rng.randint(0, 256, (2, 3, 64, 96, 3)).astype(np.uint8)).pin_memory(),
                     "jitter_f": f.pin_memory(), "jitter_i": i.pin_memory(), ("K", 0): ex[("K", 0)].pin_memory(), ("inv_K", 0): ex[("inv_K", 0)].pin_memory()})
    runner = IP.U8HostFedRunner(g, host[0])
    runner.feed(host[0])
    losses = []
    for s in range(4):
        l = runner.run()
        if s < 3:
            runner.feed(host[(s + 1) % 2])
        losses.append(float(l))
    assert all(np.isfinite(losses))
    want = torch.from_numpy(host[1]["frames_u8"].numpy()[:, 1]).permute(0, 3, 1, 2).float() / 255.0
    if int(host[1]["jitter_i"][0, 5]):
        want[0] = want[0].flip(-1)
    if int(host[1]["jitter_i"][1, 5]):
        want[1] = want[1].flip(-1)
    assert torch.equal(g.static_inputs[("color", 0, 0)].cpu(), want)      # the last batch fed is what the step saw
