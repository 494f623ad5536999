"""Golden state_dict key tables and forward outputs of the reference's D-HRNet, Lite-Mono, FusionModule and IFRNet
(run in the build container only: imports the unmodified reference from /root/reference).
Writes tests/golden/net_keys.json and tests/golden/net_*.npz, consumed by tests/test_networks_more.py."""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import ref_harness  # noqa: E402
import net_fill  # noqa: E402

T = ref_harness.import_reference(192, 640, 2)
import networks  # noqa: E402  (the reference's)
import torch  # noqa: E402

keys = {}


def dump(name, m):
    keys[name] = {k: list(v.shape) for k, v in m.state_dict().items()}


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, "net_%s.npz" % name), **{k: v.detach().numpy().astype(np.float32) for k, v in arrs.items()})


torch.manual_seed(0)
# ---- D-HRNet ---------------------------------------------------------------------------------------------------
enc = networks.DHRNet.DepthEncoder(18, False)
dec = networks.DHRNet.DepthDecoder(enc.num_ch_enc, range(1))
dump("DHRNet.DepthEncoder", enc)
dump("DHRNet.DepthDecoder", dec)
net_fill.fill_(enc), net_fill.fill_(dec)
enc.train(), dec.train()
x = net_fill.seeded_input((2, 3, 64, 96), 11)
feats = enc(x)
disp = dec(feats)[("disp", 0)]
save("dhrnet", disp=disp, f0=feats[0][:, :4], f4=feats[4][:, :8], f2=feats[2][:, :4])
# ---- Lite-Mono ---------------------------------------------------------------------------------------------------
for model in ("lite-mono", "lite-mono-small", "lite-mono-tiny", "lite-mono-8m"):
    e = networks.LiteMono.DepthEncoder(model=model, drop_path_rate=0.2, width=640, height=192)
    dump("LiteMono.DepthEncoder[%s]" % model, e)
enc = networks.LiteMono.DepthEncoder(model="lite-mono", drop_path_rate=0.2, width=640, height=192)
dec = networks.LiteMono.DepthDecoder(enc.num_ch_enc, range(1))
dump("LiteMono.DepthDecoder", dec)
net_fill.fill_(enc), net_fill.fill_(dec)
enc.eval(), dec.eval()   # eval: DropPath is stochastic in train mode
x = net_fill.seeded_input((2, 3, 64, 96), 12)
with torch.no_grad():
    feats = enc(x)
    disp = dec(feats)[("disp", 0)]
save("litemono", disp=disp, f0=feats[0][:, :4], f2=feats[2][:, :8])
# ---- FusionModule ---------------------------------------------------------------------------------------------------
for backbone, chans in (("ResNet18", [64, 64, 128, 256, 512]), ("LiteMono", [48, 80, 128])):
    args = types.SimpleNamespace(backbone=backbone)
    fm = networks.FusionModule(args, np.array(chans))
    dump("FusionModule[%s]" % backbone, fm)
    net_fill.fill_(fm)
    B, H, W = 2, 64, 96
    first = 4 if backbone == "LiteMono" else 2
    feats3 = []
    for k in range(3):
        fl = []
        for i, c in enumerate(chans):
            s = first * 2 ** i
            fl.append(net_fill.seeded_input((B, c, H // s, W // s), 100 + 10 * k + i) - 0.5)
        feats3.append(fl)
    flows = [3.0 * (net_fill.seeded_input((B, 2, H, W), 200 + k) - 0.5) for k in range(2)]
    mask = net_fill.seeded_input((B, 1, H, W), 210)
    with torch.no_grad():
        out = fm(feats3, flows, mask)
    save("fusion_%s" % backbone.lower(), **{"o%d" % i: o[:, :6] for i, o in enumerate(out)},
         **{"s%d" % i: o.double().sum().float().reshape(1) for i, o in enumerate(out)})
# ---- IFRNet ---------------------------------------------------------------------------------------------------
for scale in ("small", "large"):
    m = networks.IFRNet(scale).eval()
    dump("IFRNet[%s]" % scale, m)
    net_fill.fill_(m, scale=0.7)
    img0 = net_fill.seeded_input((2, 3, 64, 128), 31)
    img1 = net_fill.seeded_input((2, 3, 64, 128), 32)
    embt = torch.full((2, 1, 1, 1), 0.5)
    with torch.no_grad():
        pred, f0, f1, mk = m(img0, img1, embt)
        g0, g1, gm = m(img0, img1, embt, onlyFlow=True)
    assert torch.equal(f0, g0)
    save("ifrnet_%s" % scale, pred=pred[:, :, ::2, ::2], flow0=f0[:, :, ::2, ::2], flow1=f1[:, :, 1::2, 1::2], mask=mk[:, :, ::2, ::2])
json.dump(keys, open(os.path.join(HERE, "net_keys.json"), "w"), indent=0, sort_keys=True)
print({k: len(v) for k, v in keys.items()})
