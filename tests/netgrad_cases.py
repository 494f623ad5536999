"""Train-mode forward + backward cases shared by the golden generator (reference modules, CPU;
tests/golden/gen_netgrad_golden.py) and the GPU parity tests (drop-in modules on the tcgen05 path;
tests/test_networks_cuda.py).  `networks` is whichever package provides the classes."""
import types

import numpy as np
import torch

import net_fill

CASES = ["resnet18_depth", "pose", "dhrnet", "litemono", "fusion_resnet18"]
FILL_SCALE = {}
B, H, W = 2, 64, 96


def build(name, networks):
    """-> (modules, run(modules, device) -> list of output tensors)"""
    if name == "resnet18_depth":
        enc = networks.monodepth2.DepthEncoder(18, False)
        dec = networks.monodepth2.DepthDecoder(enc.num_ch_enc, range(1))

        def run(m, dev="cpu"):
            feats = m[0](net_fill.seeded_input((B, 3, H, W), 41).to(dev))
            return [m[1](feats)[("disp", 0)], feats[0], feats[4]]
        return [enc, dec], run
    if name == "pose":
        pe = networks.posenet.ResnetEncoder(18, False, num_input_images=2)
        pd = networks.posenet.PoseDecoder(pe.num_ch_enc, num_input_features=1, num_frames_to_predict_for=2)

        def run(m, dev="cpu"):
            aa, tr = m[1]([m[0](net_fill.seeded_input((B, 6, H, W), 42).to(dev))])
            return [aa, tr]
        return [pe, pd], run
    if name == "dhrnet":
        enc = networks.DHRNet.DepthEncoder(18, False)
        dec = networks.DHRNet.DepthDecoder(enc.num_ch_enc, range(1))

        def run(m, dev="cpu"):
            feats = m[0](net_fill.seeded_input((B, 3, H, W), 43).to(dev))
            return [m[1](feats)[("disp", 0)], feats[1], feats[4]]
        return [enc, dec], run
    if name == "litemono":
        enc = networks.LiteMono.DepthEncoder(model="lite-mono", drop_path_rate=0.0, width=640, height=192)
        dec = networks.LiteMono.DepthDecoder(enc.num_ch_enc, range(1))

        def run(m, dev="cpu"):
            feats = m[0](net_fill.seeded_input((B, 3, H, W), 44).to(dev))
            return [m[1](feats)[("disp", 0)], feats[0], feats[2]]
        return [enc, dec], run
    if name == "fusion_resnet18":
        chans = [64, 64, 128, 256, 512]
        fm = networks.FusionModule(types.SimpleNamespace(backbone="ResNet18"), np.array(chans))

        def run(m, dev="cpu"):
            feats3 = [[(net_fill.seeded_input((B, c, H // (2 * 2 ** i), W // (2 * 2 ** i)), 100 + 10 * k + i) - 0.5).to(dev)
                       for i, c in enumerate(chans)] for k in range(3)]
            flows = [(3.0 * (net_fill.seeded_input((B, 2, H, W), 200 + k) - 0.5)).to(dev) for k in range(2)]
            mask = net_fill.seeded_input((B, 1, H, W), 210).to(dev)
            return list(m[0](feats3, flows, mask))
        return [fm], run
    raise KeyError(name)


def loss_of(outs):
    loss = 0.0
    for i, o in enumerate(outs):
        r = (net_fill.seeded_input(tuple(o.shape), 900 + i) - 0.5).to(o.device)
        loss = loss + (o * r).sum()
    return loss


def _sub(t, cap=20000):
    f = t.detach().float().reshape(-1)
    st = max(1, f.numel() // cap)
    return f[::st].cpu().numpy().copy()


def named_params(mods):
    out = []
    for mi, m in enumerate(mods):
        for n, p in m.named_parameters():
            out.append(("%d.%s" % (mi, n), p))
    return out


def record(mods, outs):
    rec = {}
    for i, o in enumerate(outs):
        rec["out_%d" % i] = _sub(o)
        rec["out_%d_abs_sum" % i] = np.float64(o.detach().double().abs().sum().item())
    names, gsum, gabs = [], [], []
    for n, p in named_params(mods):
        names.append(n)
        g = p.grad
        gsum.append(0.0 if g is None else float(g.double().sum()))
        gabs.append(0.0 if g is None else float(g.double().abs().sum()))
    rec["names"] = np.array(names)
    rec["gsum"] = np.array(gsum, np.float64)
    rec["gabs"] = np.array(gabs, np.float64)
    with_grad = [(n, p) for n, p in named_params(mods) if p.grad is not None]
    rec["g_first"] = _sub(with_grad[0][1].grad, 50000)
    rec["g_last"] = _sub(with_grad[-1][1].grad, 50000)
    rec["g_first_name"], rec["g_last_name"] = np.array(with_grad[0][0]), np.array(with_grad[-1][0])
    for m in mods:
        for mod in m.modules():
            if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
                rec["bn_mean"] = mod.running_mean.detach().float().cpu().numpy()
                rec["bn_var"] = mod.running_var.detach().float().cpu().numpy()
                return rec
    return rec
