"""GPU parity of the stand-alone layers.py drop-ins (through the C ABI) against the CPU oracle, the golden vectors
of the unmodified reference, and -- for gradients -- a plain-torch fp64 restatement of the same op on the GPU."""
import os

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _env():
    import torch
    from mono_vifi_b200 import layers as L
    from oracle import f1_oracle as O
    return torch, L, O, torch.device("cuda:0")


def load(name):
    g = np.load(os.path.join(GOLD, "f1_%s.npz" % name))
    seed, B, H, W, structured, with_mask, f0, f1, f2 = [int(x) for x in g["spec"]]
    c = synth.make_case(seed, B, H, W, bool(structured), bool(with_mask), 1 if f1 else 2)
    return g, c, (B, H, W)


@pytest.mark.parametrize("name", ["tiny", "small", "odd"])
def test_geometry_chain_bit_exact(name):
    torch, L, O, dev = _env()
    g, c, (B, H, W) = load(name)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sd, depth = L.disp_to_depth(t(c["disp"]), 0.1, 100.0)
    assert np.array_equal(depth.cpu().numpy(), g["depth"])
    bp, pj = L.BackprojectDepth(B, H, W).to(dev), L.Project3D(B, H, W).to(dev)
    cam = bp(depth, t(c["inv_K"]))
    assert np.array_equal(cam.cpu().numpy(), O.backproject(g["depth"], c["inv_K"]))
    for k in range(2):
        pix = pj(cam, t(c["K"]), t(g["T"][k]))
        assert pix.shape == (B, H, W, 2)
        assert np.array_equal(pix.cpu().numpy(), g["grid"][k]), "sampling grid differs from the reference's bits"


@pytest.mark.parametrize("name", ["tiny", "small", "odd"])
def test_ssim_smooth_silog_forward(name):
    torch, L, O, dev = _env()
    g, c, (B, H, W) = load(name)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    s = L.SSIM()(t(g["warp"][0]), t(c["tgt"])).cpu().numpy()
    np.testing.assert_allclose(s, g["ssim0"], rtol=0, atol=3e-4)          # fp32 E[x^2]-mu^2 cancellation
    np.testing.assert_allclose(s, O.ssim(g["warp"][0], c["tgt"]), rtol=0, atol=3e-4)
    mean = c["disp"].mean(axis=(2, 3), keepdims=True, dtype=np.float64).astype(np.float32)
    nd = c["disp"] / (mean + np.float32(1e-7))
    sm = float(L.get_smooth_loss(t(nd), t(c["tgt"])))
    assert abs(sm - float(g["smooth"])) <= 1e-5 * float(g["smooth"])
    noise1 = c["noise"][:, :1]
    depth = g["depth"]
    depth2 = (depth * (np.float32(1.0) + np.float32(0.1) * np.clip(noise1, -3, 3) * np.float32(0.3))).astype(np.float32)
    v = float(L.si_log_depth_loss(t(depth), t(depth2)))
    assert abs(v - float(g["si_log"])) <= 1e-4 * abs(float(g["si_log"])) + 1e-9
    m = (noise1 > 0).astype(np.float32)
    v = float(L.si_log_depth_loss(t(depth), t(depth2), t(m)))
    assert abs(v - float(g["si_log_masked"])) <= 1e-4 * abs(float(g["si_log_masked"])) + 1e-9


def _ref_ssim(x, y):
    import torch.nn.functional as F
    x, y = F.pad(x, (1, 1, 1, 1), mode="reflect"), F.pad(y, (1, 1, 1, 1), mode="reflect")
    p = lambda a: F.avg_pool2d(a, 3, 1)
    mx, my = p(x), p(y)
    sx, sy, sxy = p(x * x) - mx * mx, p(y * y) - my * my, p(x * y) - mx * my
    n = (2 * mx * my + 0.01 ** 2) * (2 * sxy + 0.03 ** 2)
    d = (mx * mx + my * my + 0.01 ** 2) * (sx + sy + 0.03 ** 2)
    return ((1 - n / d) / 2).clamp(0, 1)


def _ref_smooth(disp, img):
    import torch
    gx = (disp[:, :, :, :-1] - disp[:, :, :, 1:]).abs() * torch.exp(-(img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, True))
    gy = (disp[:, :, :-1, :] - disp[:, :, 1:, :]).abs() * torch.exp(-(img[:, :, :-1, :] - img[:, :, 1:, :]).abs().mean(1, True))
    return gx.mean() + gy.mean()


def _ref_silog(pred, target, mask, beta=0.5):
    import torch
    mask = torch.ones_like(pred) if mask is None else mask
    mask = mask[:, 0]
    ld = torch.log(pred[:, 0] + 1e-7) * mask - torch.log(target[:, 0] + 1e-7) * mask
    n = mask.sum(1).sum(1) + 1e-8
    return ((ld ** 2).sum(1).sum(1) / n - beta * ld.sum(1).sum(1) ** 2 / n ** 2).mean()


@pytest.mark.parametrize("shape", [(2, 3, 16, 24), (1, 3, 21, 35), (2, 3, 64, 96)])
def test_gradients_vs_torch_fp64(shape):
    torch, L, O, dev = _env()
    B, C, H, W = shape
    gen = torch.Generator(device="cpu").manual_seed(H * W)
    r = lambda *s: torch.rand(*s, generator=gen).to(dev)
    # SSIM: gradients to both arguments
    x, y, w = r(B, C, H, W).requires_grad_(True), r(B, C, H, W).requires_grad_(True), r(B, C, H, W)
    (L.SSIM()(x, y) * w).sum().backward()
    xd, yd = x.detach().double().requires_grad_(True), y.detach().double().requires_grad_(True)
    (_ref_ssim(xd, yd) * w.double()).sum().backward()
    for a, b in ((x.grad, xd.grad), (y.grad, yd.grad)):
        sc = b.abs().max()
        bad = ((a.double() - b).abs() > 2e-3 * sc + 2e-3 * b.abs()).float().mean()
        assert bad <= 2e-3, float(bad)   # isolated windows sit on the clamp boundary / lose bits to cancellation
    # smoothness
    d, im = r(B, 1, H, W).requires_grad_(True), r(B, 3, H, W)
    (L.get_smooth_loss(d, im) * 1.7).backward()
    dd = d.detach().double().requires_grad_(True)
    (_ref_smooth(dd, im.double()) * 1.7).backward()
    np.testing.assert_allclose(d.grad.cpu().numpy(), dd.grad.cpu().numpy(), rtol=1e-4, atol=1e-9)
    # SI-log, masked and unmasked, gradients to both depths
    for use_mask in (False, True):
        p, q = (r(B, 1, H, W) * 9 + 0.5).requires_grad_(True), (r(B, 1, H, W) * 9 + 0.5).requires_grad_(True)
        m = (r(B, 1, H, W) > 0.3).float() if use_mask else None
        L.si_log_depth_loss(p, q, m).backward()
        pd, qd = p.detach().double().requires_grad_(True), q.detach().double().requires_grad_(True)
        _ref_silog(pd, qd, None if m is None else m.double()).backward()
        np.testing.assert_allclose(p.grad.cpu().numpy(), pd.grad.cpu().numpy(), rtol=2e-3, atol=1e-9)
        np.testing.assert_allclose(q.grad.cpu().numpy(), qd.grad.cpu().numpy(), rtol=2e-3, atol=1e-9)
    # geometry chain: disp -> depth -> cam points -> pix coords, gradients to disp and T
    c = synth.make_case(3, B, H, W)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    from tests_helpers import synth_T
    T0 = t(synth_T(c["axisangle"][0], c["translation"][0], False))
    disp, T = t(c["disp"]).requires_grad_(True), T0.clone().requires_grad_(True)
    wgt = r(B, H, W, 2)
    _, depth = L.disp_to_depth(disp, 0.1, 100.0)
    pix = L.Project3D(B, H, W).to(dev)(L.BackprojectDepth(B, H, W).to(dev)(depth, t(c["inv_K"])), t(c["K"]), T)
    (pix * wgt).sum().backward()
    dd, Td = disp.detach().double().requires_grad_(True), T0.double().requires_grad_(True)
    depth_d = 1 / (0.01 + 9.99 * dd)
    v, u = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float64), torch.arange(W, device=dev, dtype=torch.float64), indexing="ij")
    pixc = torch.stack([u.reshape(-1), v.reshape(-1), torch.ones(H * W, device=dev, dtype=torch.float64)], 0)[None].repeat(B, 1, 1)
    cam = torch.matmul(t(c["inv_K"]).double()[:, :3, :3], pixc) * depth_d.view(B, 1, -1)
    cam = torch.cat([cam, torch.ones(B, 1, H * W, device=dev, dtype=torch.float64)], 1)
    Pm = torch.matmul(t(c["K"]).double(), Td)[:, :3]
    cp = torch.matmul(Pm, cam)
    pc = (cp[:, :2] / (cp[:, 2:3] + 1e-7)).view(B, 2, H, W).permute(0, 2, 3, 1)
    pc = torch.stack([pc[..., 0] / (W - 1), pc[..., 1] / (H - 1)], -1)
    pc = (pc - 0.5) * 2
    (pc * wgt.double()).sum().backward()
    np.testing.assert_allclose(disp.grad.cpu().numpy(), dd.grad.cpu().numpy(), rtol=2e-3, atol=1e-6 * float(dd.grad.abs().max()))
    np.testing.assert_allclose(T.grad.cpu().numpy(), Td.grad.cpu().numpy(), rtol=2e-3, atol=2e-4 * float(Td.grad.abs().max()))


def test_cpu_tensors_are_rejected():
    torch, L, O, dev = _env()
    with pytest.raises(RuntimeError):
        L.SSIM()(torch.rand(1, 3, 8, 8), torch.rand(1, 3, 8, 8))
