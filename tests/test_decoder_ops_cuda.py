"""GPU parity of the fused upsample + concat + reflection-pad kernels and of the ELU-in-epilogue convolution against the
reference's op sequence (F.interpolate + torch.cat + nn.ReflectionPad2d + conv + ELU: monodepth2.py:86-93, layers.py:106-139)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [(2, 8, 0, 6, 10, False), (2, 16, 8, 8, 12, True), (1, 32, 64, 12, 40, True), (3, 4, 4, 4, 4, False),
                                  (1, 256, 0, 6, 20, False)])
def test_upcat_pad_forward_backward_exact(case):
    import torch
    import torch.nn.functional as F
    from mono_vifi_b200 import decoder_ops
    B, Ca, Cs, H, W, up = case
    g = torch.Generator(device="cuda").manual_seed(3)
    Ha, Wa = (H // 2, W // 2) if up else (H, W)
    a = torch.randn(B, Ca, Ha, Wa, device="cuda", generator=g, requires_grad=True)
    skip = torch.randn(B, Cs, H, W, device="cuda", generator=g, requires_grad=True) if Cs else None
    y = decoder_ops.upcat_pad(a, skip, up)
    ar = a.detach().clone().requires_grad_(True)
    sr = skip.detach().clone().requires_grad_(True) if Cs else None
    x = F.interpolate(ar, scale_factor=2, mode="nearest") if up else ar
    if Cs:
        x = torch.cat([x, sr], 1)
    ref = F.pad(x, (1, 1, 1, 1), mode="reflect")
    assert torch.equal(y, ref)                       # pure data movement: bit-exact
    gy = torch.randn(ref.shape, device="cuda", generator=g)
    y.backward(gy)
    ref.backward(gy)
    assert torch.allclose(a.grad, ar.grad, rtol=1e-6, atol=1e-6)
    if Cs:
        assert torch.allclose(skip.grad, sr.grad, rtol=1e-6, atol=1e-6)


def test_convblock_fast_path_matches_reference_sequence():
    import torch
    import torch.nn.functional as F
    from mono_vifi_b200 import conv, layers
    conv.set_backend("tcgen05")
    torch.manual_seed(0)
    blk = layers.ConvBlock(48, 32).cuda()
    x = torch.randn(2, 16, 8, 12, device="cuda", requires_grad=True)
    skip = torch.randn(2, 32, 16, 24, device="cuda", requires_grad=True)
    y = blk.forward_upcat(x, skip, upsample=True)
    xr, sr = x.detach().double().requires_grad_(True), skip.detach().double().requires_grad_(True)
    w, b = blk.conv.conv.weight.detach().double().requires_grad_(True), blk.conv.conv.bias.detach().double().requires_grad_(True)
    z = torch.cat([F.interpolate(xr, scale_factor=2, mode="nearest"), sr], 1)
    ref = F.elu(F.conv2d(F.pad(z, (1, 1, 1, 1), mode="reflect"), w, b))
    assert (y.double() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    gy = torch.randn(ref.shape, device="cuda")
    y.backward(gy)
    ref.backward(gy.double())
    for got, want in ((x.grad, xr.grad), (skip.grad, sr.grad), (blk.conv.conv.weight.grad, w.grad), (blk.conv.conv.bias.grad, b.grad)):
        assert (got.double() - want).abs().max().item() <= 4e-3 * want.abs().max().item()


@pytest.mark.parametrize("shape", [(2, 8, 6, 10), (1, 64, 32, 48), (3, 4, 7, 9), (2, 64, 96, 320)])
def test_maxpool3s2_matches_torch(shape):
    import torch
    import torch.nn.functional as F
    from mono_vifi_b200 import decoder_ops
    g = torch.Generator(device="cuda").manual_seed(5)
    # quantised values: plenty of ties inside windows, so the first-maximum rule is exercised
    x = (torch.randint(0, 6, shape, device="cuda", generator=g).float() / 4).requires_grad_(True)
    y = decoder_ops.maxpool3s2(x)
    xr = x.detach().clone().requires_grad_(True)
    ref = F.max_pool2d(xr, 3, 2, 1)
    assert torch.equal(y, ref)
    gy = torch.randn(ref.shape, device="cuda", generator=g)
    y.backward(gy)
    ref.backward(gy)
    assert torch.allclose(x.grad, xr.grad, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("case", [(2, 64, 12, 20, True, True), (3, 128, 6, 10, False, True), (2, 16, 9, 7, True, False),
                                  (12, 64, 96, 320, True, True), (1, 512, 6, 20, False, True), (2, 144, 6, 10, True, True),
                                  (2, 18, 24, 40, True, True), (3, 18, 9, 14, False, False), (12, 18, 48, 160, True, True)])   # HRNet: C % 4 == 2
def test_fused_bn_add_relu_matches_torch(case):
    """relu(bn(x) + identity) in training mode: output, running statistics and all gradients against torch in fp64"""
    import torch
    from mono_vifi_b200 import bn_act
    B, C, H, W, with_id, relu = case
    g = torch.Generator(device="cuda").manual_seed(9)
    x = (2.0 * torch.randn(B, C, H, W, device="cuda", generator=g) + 0.5).requires_grad_(True)
    idn = torch.randn(B, C, H, W, device="cuda", generator=g).requires_grad_(True) if with_id else None
    bn = torch.nn.BatchNorm2d(C).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.3)
    ref_bn = torch.nn.BatchNorm2d(C).cuda().double().train()
    ref_bn.load_state_dict({k: (v.double() if v.dtype.is_floating_point else v) for k, v in bn.state_dict().items()})
    before = dict(bn_act.launches)
    y = bn_act.bn_act(bn, x, idn, relu)
    assert bn_act.launches["bn_fwd"] == before["bn_fwd"] + 1
    xr = x.detach().double().requires_grad_(True)
    ir = idn.detach().double().requires_grad_(True) if with_id else None
    yr = ref_bn(xr)
    if with_id:
        yr = yr + ir
    if relu:
        yr = torch.relu(yr)
    assert (y.double() - yr).abs().max().item() <= 2e-5 * max(1.0, yr.abs().max().item())
    assert torch.allclose(bn.running_mean.double(), ref_bn.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(bn.running_var.double(), ref_bn.running_var, rtol=1e-5, atol=1e-6)
    assert int(bn.num_batches_tracked) == int(ref_bn.num_batches_tracked) == 1
    gy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(gy)
    yr.backward(gy.double())
    pairs = [(x.grad, xr.grad), (bn.weight.grad, ref_bn.weight.grad), (bn.bias.grad, ref_bn.bias.grad)]
    if with_id:
        pairs.append((idn.grad, ir.grad))
    for got, want in pairs:
        assert (got.double() - want).abs().max().item() <= 1e-4 * max(1e-3, want.abs().max().item())


@pytest.mark.parametrize("case", [(2, 16, 9, 13, 2, True), (12, 64, 24, 80, 1, True), (1, 256, 6, 20, 0, True), (3, 32, 7, 5, 2, False),
                                  (2, 1024, 3, 4, 1, True)])
def test_act_backward_and_bias_gradient_match_torch(case):
    """mvf_act_bwd_bias vs aten elu_backward / threshold_backward + sum((0,2,3)) (the backward of layers.py:68-117)"""
    import torch
    from mono_vifi_b200 import conv_tc
    B, C, H, W, act, want_bias = case
    g = torch.Generator(device="cuda").manual_seed(11)
    gy = torch.randn(B, C, H, W, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    y = torch.randn(B, C, H, W, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    if act == 2:
        y = torch.where(y > 0, y, torch.expm1(y))
    elif act == 1:
        y = torch.relu(y)
    gpre, gb = conv_tc.act_bwd_bias(gy, y if act else None, act, want_bias)
    if act == 1:
        ref = torch.ops.aten.threshold_backward(gy, y, 0.0)
    elif act == 2:
        ref = torch.ops.aten.elu_backward(gy, 1.0, 1.0, 1.0, True, y)
    else:
        ref = gy
    assert torch.equal(gpre, ref) or torch.allclose(gpre, ref, rtol=1e-6, atol=1e-7)
    if want_bias:
        refb = ref.double().sum((0, 2, 3))
        assert torch.allclose(gb.double(), refb, rtol=1e-5, atol=1e-5 * float(ref.abs().sum((0, 2, 3)).max()))
    else:
        assert gb is None
