"""Golden losses of the reference's own Trainer.process_batch (train.py:698-885, without and with the affine branch) on CPU for seeded
inputs and name-keyed weights (run in the build container only).  Writes tests/golden/step_golden.json, consumed by
tests/test_trainstep_cuda.py::test_multi_frame_step_matches_reference."""
import copy
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.append(os.path.dirname(os.path.dirname(HERE)))  # repo root: mono_vifi_b200.affine builds the affine INPUTS
import ref_harness  # noqa: E402
import net_fill  # noqa: E402

B, H, W = 2, 64, 96
T = ref_harness.import_reference(H, W, B)
import networks  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402


def inputs_for(B, H, W):
    inp = {}
    for i, f in enumerate((-1, 0, 1)):
        inp[("color", f, 0)] = net_fill.seeded_input((B, 3, H, W), 500 + i)
        inp[("color_aug", f, 0)] = net_fill.seeded_input((B, 3, H, W), 510 + i)
    K = np.array([[0.58 * W, 0, 0.5 * W, 0], [0, 1.92 * H, 0.5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    inp[("K", 0)] = torch.from_numpy(np.repeat(K[None], B, 0).copy())
    inp[("inv_K", 0)] = torch.from_numpy(np.repeat(np.linalg.pinv(K)[None], B, 0).astype(np.float32).copy())
    return inp


import affine_inputs  # noqa: E402

out = {"B": B, "H": H, "W": W}
for backbone in ("ResNet18", "DHRNet", "ResNet18_affine"):
    use_affine = backbone.endswith("_affine")     # the same step with the affine-augmentation branch (train.py:815-883)
    key, backbone = backbone, backbone.split("_")[0]
    tr = ref_harness.make_trainer(T, B, H, W)
    tr.opt.use_affine = use_affine
    tr.opt.backbone = backbone
    tr.opt.fuse_model_type = "shared_encoder"
    torch.manual_seed(0)
    m = {}
    if backbone == "ResNet18":
        m["encoder"] = networks.monodepth2.DepthEncoder(18, False)
        m["depth"] = networks.monodepth2.DepthDecoder(m["encoder"].num_ch_enc, range(1))
    else:
        m["encoder"] = networks.DHRNet.DepthEncoder(18, False)
        m["depth"] = networks.DHRNet.DepthDecoder(m["encoder"].num_ch_enc, range(1))
    m["encoder_mf"] = m["encoder"]
    m["depth_mf"] = copy.deepcopy(m["depth"])
    m["fusion_module"] = networks.FusionModule(tr.opt, m["encoder_mf"].num_ch_enc)
    m["pose_encoder"] = networks.posenet.ResnetEncoder(18, False, num_input_images=2)
    m["pose"] = networks.posenet.PoseDecoder(m["pose_encoder"].num_ch_enc, num_input_features=1, num_frames_to_predict_for=2)
    for name, mod in m.items():
        if name == "encoder_mf":
            continue
        net_fill.fill_(mod, scale=0.5 if name != "depth_mf" else 0.6)
        mod.train()
    tr.models = m
    tr.model_vfi_train = net_fill.fill_(networks.IFRNet("small"), scale=0.7).eval()
    torch.manual_seed(1)
    inp = inputs_for(B, H, W)
    if use_affine:
        affine_inputs.add_affine_inputs(inp, B, H, W)
    _, losses = tr.process_batch(inp)
    out[key] = {k: float(v) for k, v in losses.items()}
    print(key, out[key])
    # every parameter gradient of the step (reference autograd): sum and abs-sum per tensor
    losses["loss"].backward()
    names, gsum, gabs = [], [], []
    for mname, mod in m.items():
        if mname == "encoder_mf":
            continue
        for pname, p in mod.named_parameters():
            names.append("%s.%s" % (mname, pname))
            gsum.append(0.0 if p.grad is None else float(p.grad.double().sum()))
            gabs.append(0.0 if p.grad is None else float(p.grad.double().abs().sum()))
    np.savez_compressed(os.path.join(HERE, "step_grads_%s.npz" % key), names=np.array(names), gsum=np.array(gsum), gabs=np.array(gabs))
json.dump(out, open(os.path.join(HERE, "step_golden.json"), "w"), indent=1, sort_keys=True)
