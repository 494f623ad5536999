"""Host-side parity (CPU) of the drop-in D-HRNet, Lite-Mono, FusionModule and IFRNet with the unmodified reference:
state_dict key / shape tables and forward outputs on seeded inputs with identically filled weights
(fixtures: tests/golden/net_keys.json, net_*.npz written by tests/golden/gen_net_golden.py in the build container)."""
import json
import os
import types

import numpy as np
import pytest
import torch

import net_fill

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEYS = json.load(open(os.path.join(GOLD, "net_keys.json")))


def _sd(m):
    return {k: list(v.shape) for k, v in m.state_dict().items()}


def _gold(name):
    return np.load(os.path.join(GOLD, "net_%s.npz" % name))


def _close(got, want, tol=2e-5):
    want = torch.from_numpy(want)
    scale = max(1e-6, float(want.abs().max()))
    err = float((got.detach().float() - want).abs().max())
    assert err <= tol * scale + 1e-6, (err, scale)


def test_dhrnet_matches_reference():
    from mono_vifi_b200.networks import DHRNet
    torch.manual_seed(0)
    enc = DHRNet.DepthEncoder(18, False)
    dec = DHRNet.DepthDecoder(enc.num_ch_enc, range(1))
    assert _sd(enc) == KEYS["DHRNet.DepthEncoder"] and _sd(dec) == KEYS["DHRNet.DepthDecoder"]
    assert list(enc.num_ch_enc) == [64, 18, 36, 72, 144]
    net_fill.fill_(enc), net_fill.fill_(dec)
    enc.train(), dec.train()
    g = _gold("dhrnet")
    feats = enc(net_fill.seeded_input((2, 3, 64, 96), 11))
    assert enc.features is feats and [f.shape[1] for f in feats] == [64, 18, 36, 72, 144]
    out = dec(feats)
    assert list(out.keys()) == [("disp", 0)]
    _close(feats[0][:, :4], g["f0"], 1e-4), _close(feats[2][:, :4], g["f2"], 1e-4), _close(feats[4][:, :8], g["f4"], 1e-4)
    _close(out[("disp", 0)], g["disp"], 1e-4)
    with pytest.raises(AssertionError):
        DHRNet.DepthEncoder(50, False)


def test_litemono_matches_reference():
    from mono_vifi_b200.networks import LiteMono
    for model in ("lite-mono", "lite-mono-small", "lite-mono-tiny", "lite-mono-8m"):
        e = LiteMono.DepthEncoder(model=model, drop_path_rate=0.2, width=640, height=192)
        assert _sd(e) == KEYS["LiteMono.DepthEncoder[%s]" % model], model
    torch.manual_seed(0)
    enc = LiteMono.DepthEncoder(model="lite-mono", drop_path_rate=0.2, width=640, height=192)
    dec = LiteMono.DepthDecoder(enc.num_ch_enc, range(1))
    assert _sd(dec) == KEYS["LiteMono.DepthDecoder"]
    net_fill.fill_(enc), net_fill.fill_(dec)
    enc.eval(), dec.eval()
    g = _gold("litemono")
    with torch.no_grad():
        feats = enc(net_fill.seeded_input((2, 3, 64, 96), 12))
        disp = dec(feats)[("disp", 0)]
    assert [f.shape[1] for f in feats] == [48, 80, 128] and disp.shape == (2, 1, 64, 96)
    _close(feats[0][:, :4], g["f0"], 1e-4), _close(feats[2][:, :8], g["f2"], 1e-4), _close(disp, g["disp"], 1e-4)
    # 320x1024 uses the wider dilation schedule
    assert LiteMono.DepthEncoder(model="lite-mono", width=1024, height=320).dilation[2][-1] == 10


@pytest.mark.parametrize("backbone,chans", [("ResNet18", [64, 64, 128, 256, 512]), ("LiteMono", [48, 80, 128])])
def test_fusion_module_matches_reference(backbone, chans):
    from mono_vifi_b200.networks import FusionModule
    fm = FusionModule(types.SimpleNamespace(backbone=backbone), np.array(chans))
    assert _sd(fm) == KEYS["FusionModule[%s]" % backbone]
    net_fill.fill_(fm)
    B, H, W = 2, 64, 96
    first = 4 if backbone == "LiteMono" else 2
    feats3 = [[net_fill.seeded_input((B, c, H // (first * 2 ** i), W // (first * 2 ** i)), 100 + 10 * k + i) - 0.5
               for i, c in enumerate(chans)] for k in range(3)]
    flows = [3.0 * (net_fill.seeded_input((B, 2, H, W), 200 + k) - 0.5) for k in range(2)]
    mask = net_fill.seeded_input((B, 1, H, W), 210)
    with torch.no_grad():
        out = fm(feats3, flows, mask)
    g = _gold("fusion_%s" % backbone.lower())
    for i, o in enumerate(out):
        assert o.shape[1] == chans[i]
        _close(o[:, :6], g["o%d" % i], 1e-4)
        assert abs(float(o.double().sum()) - float(g["s%d" % i][0])) <= 1e-3 * max(1.0, abs(float(g["s%d" % i][0])))


@pytest.mark.parametrize("scale", ["small", "large"])
def test_ifrnet_matches_reference(scale):
    from mono_vifi_b200 import networks as N
    m = N.IFRNet(scale).eval()
    assert _sd(m) == KEYS["IFRNet[%s]" % scale]
    net_fill.fill_(m, scale=0.7)
    img0, img1 = net_fill.seeded_input((2, 3, 64, 128), 31), net_fill.seeded_input((2, 3, 64, 128), 32)
    embt = torch.full((2, 1, 1, 1), 0.5)
    with torch.no_grad():
        pred, f0, f1, mk = m(img0, img1, embt)
        g0, g1, gm = m(img0, img1, embt, onlyFlow=True)
    assert torch.equal(f0, g0) and torch.equal(mk, gm)
    g = _gold("ifrnet_%s" % scale)
    _close(pred[:, :, ::2, ::2], g["pred"], 1e-4), _close(f0[:, :, ::2, ::2], g["flow0"], 1e-4)
    _close(f1[:, :, 1::2, 1::2], g["flow1"], 1e-4), _close(mk[:, :, ::2, ::2], g["mask"], 1e-4)
    with pytest.raises(NotImplementedError):
        m(img0, img1, embt, imgt=img0)


@pytest.mark.parametrize("name", ["resnet18_depth", "pose", "dhrnet", "litemono", "fusion_resnet18"])
def test_train_gradients_match_reference_on_cpu(name):
    """Module wiring of forward AND backward (train mode) against the reference's autograd, on the host (torch ops):
    the GPU twin (tests/test_networks_cuda.py) holds the tcgen05 path to the same fixtures."""
    import netgrad_cases as NC
    from mono_vifi_b200 import networks
    g = np.load(os.path.join(GOLD, "netgrad_%s.npz" % name))
    torch.manual_seed(0)
    mods, run = NC.build(name, networks)
    for m in mods:
        net_fill.fill_(m)
        m.train()
    outs = run(mods)
    NC.loss_of(outs).backward()
    rec = NC.record(mods, outs)
    for i in range(len(outs)):
        scale = max(1e-6, float(np.abs(g["out_%d" % i]).max()))
        assert float(np.abs(rec["out_%d" % i] - g["out_%d" % i]).max()) <= 1e-4 * scale
    assert list(rec["names"]) == list(g["names"])
    assert np.all(np.abs(rec["gsum"] - g["gsum"]) <= 2e-4 * g["gabs"] + 1e-7)
    assert np.all(np.abs(rec["gabs"] - g["gabs"]) <= 2e-4 * g["gabs"] + 1e-7)
    for k in ("g_first", "g_last"):
        assert float(np.abs(rec[k] - g[k]).max()) <= 2e-4 * max(1e-9, float(np.abs(g[k]).max()))
