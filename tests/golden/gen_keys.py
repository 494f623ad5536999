"""Dump the state_dict key -> shape tables of the reference's networks (run in the build container only).
Output: tests/golden/state_dict_keys.json, consumed by tests/test_networks_api.py."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402

T = ref_harness.import_reference(192, 640, 2)
import networks  # noqa: E402  (the reference's)
import torch  # noqa: E402

out = {}


def dump(name, m):
    out[name] = {k: list(v.shape) for k, v in m.state_dict().items()}


enc = networks.monodepth2.DepthEncoder(18, False)
dump("monodepth2.DepthEncoder18", enc)
dump("monodepth2.DepthDecoder18", networks.monodepth2.DepthDecoder(enc.num_ch_enc, range(1)))
enc50 = networks.monodepth2.DepthEncoder(50, False)
dump("monodepth2.DepthEncoder50", enc50)
dump("monodepth2.DepthDecoder50", networks.monodepth2.DepthDecoder(enc50.num_ch_enc, range(4)))
pe = networks.posenet.ResnetEncoder(18, False, num_input_images=2)
dump("posenet.ResnetEncoder18x2", pe)
dump("posenet.PoseDecoder", networks.posenet.PoseDecoder(pe.num_ch_enc, num_input_features=1, num_frames_to_predict_for=2))
json.dump(out, open(os.path.join(HERE, "state_dict_keys.json"), "w"), indent=0, sort_keys=True)
print({k: len(v) for k, v in out.items()})
