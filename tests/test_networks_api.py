"""Host-side checks (CPU): the drop-in networks expose the reference's module API -- state_dict keys and shapes
(fixture generated from the unmodified reference by tests/golden/gen_keys.py), attributes, output structure --
and the C-ABI library exports every symbol include/monovifi_b200.h declares."""
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))


def _sd(m):
    return {k: list(v.shape) for k, v in m.state_dict().items()}


def test_state_dict_keys_match_reference():
    from mono_vifi_b200 import networks as N
    enc = N.monodepth2.DepthEncoder(18, False)
    assert _sd(enc) == KEYS["monodepth2.DepthEncoder18"]
    assert _sd(N.monodepth2.DepthDecoder(enc.num_ch_enc, range(1))) == KEYS["monodepth2.DepthDecoder18"]
    enc50 = N.monodepth2.DepthEncoder(50, False)
    assert _sd(enc50) == KEYS["monodepth2.DepthEncoder50"]
    assert _sd(N.monodepth2.DepthDecoder(enc50.num_ch_enc, range(4))) == KEYS["monodepth2.DepthDecoder50"]
    pe = N.posenet.ResnetEncoder(18, False, num_input_images=2)
    assert _sd(pe) == KEYS["posenet.ResnetEncoder18x2"]
    assert _sd(N.posenet.PoseDecoder(pe.num_ch_enc, num_input_features=1, num_frames_to_predict_for=2)) == KEYS["posenet.PoseDecoder"]


def test_forward_structure_cpu():
    """Networks are torch modules over Conv2d; with the cuDNN/ATen conv backend they also run on CPU tensors."""
    from mono_vifi_b200 import networks as N
    torch.manual_seed(0)
    enc = N.monodepth2.DepthEncoder(18, False)
    dec = N.monodepth2.DepthDecoder(enc.num_ch_enc, range(1))
    x = torch.rand(1, 3, 64, 96)
    feats = enc(x)
    assert [f.shape[1] for f in feats] == [64, 64, 128, 256, 512] and feats[0].shape[2:] == (32, 48)
    assert enc.features is feats
    out = dec(feats)
    assert list(out.keys()) == [("disp", 0)] and out[("disp", 0)].shape == (1, 1, 64, 96)
    assert float(out[("disp", 0)].min()) >= 0 and float(out[("disp", 0)].max()) <= 1
    pe = N.posenet.ResnetEncoder(18, False, num_input_images=2)
    pd = N.posenet.PoseDecoder(pe.num_ch_enc, num_input_features=1, num_frames_to_predict_for=2)
    aa, tr = pd([pe(torch.rand(2, 6, 64, 96))])
    assert aa.shape == (2, 2, 1, 3) and tr.shape == (2, 2, 1, 3)
    with pytest.raises(ValueError):
        N.monodepth2.DepthEncoder(19, False)


def test_transformation_from_parameters_matches_golden():
    from mono_vifi_b200 import layers as L
    import synth
    g = np.load(os.path.join(ROOT, "tests", "golden", "f1_small.npz"))
    c = synth.make_case(12, 2, 64, 96, True, True, 2)
    for k in range(2):
        T = L.transformation_from_parameters(torch.from_numpy(c["axisangle"][k]), torch.from_numpy(c["translation"][k]),
                                             invert=(k == 1))
        assert np.array_equal(T.numpy(), g["T"][k])   # same op sequence -> same bits
        P = L.matmul_KT(torch.from_numpy(c["K"]), T)[:, :3]
        assert np.array_equal(P.numpy(), g["P"][k])   # layers.py:212 (K @ T), the reference's bits


def test_library_exports_every_declared_symbol():
    from mono_vifi_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "monovifi_b200.h")).read()
    declared = set(re.findall(r"\b(mvf_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mvf_status"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = _lib.lib()   # raises if the .so is missing or lacks a symbol
    for name in declared:
        assert hasattr(lib, name)
    assert lib.mvf_version() >= 100
    assert lib.mvf_f1_workspace_bytes(12) >= 12 * 28 * 8


def test_bad_arguments_return_error_codes_without_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised on a CPU-only host."""
    from mono_vifi_b200 import _lib
    lib = _lib.lib()
    p = _lib.f1_params(0, 8, 8)
    rc = lib.mvf_f1_forward(p, *([None] * 16), None, 0, None)
    assert rc == -1 and b"bad params" in lib.mvf_last_error()
    p = _lib.f1_params(1, 2, 8)
    assert lib.mvf_f1_backward(p, *([None] * 14), None, 0, None) == -1
    assert lib.mvf_ssim_fwd(None, None, None, 1, 8, 8, None) == -1
