"""GPU eager-PyTorch arm of the benchmark ("the kernel to beat"): the single-frame step of
BASELINE configs[1] written with nothing but stock torch / torchvision modules -- cuDNN convolutions, nn.BatchNorm2d,
F.grid_sample, nn.AvgPool2d SSIM, torch.optim.AdamW, one kernel launch per op, no CUDA graph, none of this repository's
kernels.  It follows what the reference executes on a GPU (networks/monodepth2.py:11-96, networks/posenet.py:10-137,
layers.py:16-25, 168-222, 231-242, 261-290, train.py:956-1051, 659-666); the reference itself is not present on the GPU
box, so this is a restatement used ONLY as a timed baseline by bench.py (never by the product path)."""
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision


class ConvBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.pad, self.conv, self.act = nn.ReflectionPad2d(1), nn.Conv2d(int(cin), int(cout), 3), nn.ELU(inplace=True)

    def forward(self, x):
        return self.act(self.conv(self.pad(x)))


class DepthNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.enc = torchvision.models.resnet18(weights=None)
        ce, cd = [64, 64, 128, 256, 512], [16, 32, 64, 128, 256]
        self.up0 = nn.ModuleList([ConvBlock(ce[-1] if i == 4 else cd[i + 1], cd[i]) for i in range(5)])
        self.up1 = nn.ModuleList([ConvBlock(cd[i] + (ce[i - 1] if i > 0 else 0), cd[i]) for i in range(5)])
        self.disp = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(cd[0], 1, 3))

    def forward(self, img):
        e = self.enc
        f = [e.relu(e.bn1(e.conv1((img - 0.45) / 0.225)))]
        f.append(e.layer1(e.maxpool(f[-1])))
        f.append(e.layer2(f[-1]))
        f.append(e.layer3(f[-1]))
        f.append(e.layer4(f[-1]))
        x = f[-1]
        for i in range(4, -1, -1):
            x = F.interpolate(self.up0[i](x), scale_factor=2, mode="nearest")
            if i > 0:
                x = torch.cat([x, f[i - 1]], 1)
            x = self.up1[i](x)
        return torch.sigmoid(self.disp(x))


class PoseNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.enc = torchvision.models.resnet18(weights=None)
        self.enc.conv1 = nn.Conv2d(6, 64, 7, 2, 3, bias=False)
        self.squeeze = nn.Conv2d(512, 256, 1)
        self.p0, self.p1, self.p2 = nn.Conv2d(256, 256, 3, 1, 1), nn.Conv2d(256, 256, 3, 1, 1), nn.Conv2d(256, 12, 1)

    def forward(self, a, b):
        e = self.enc
        x = e.relu(e.bn1(e.conv1((torch.cat([a, b], 1) - 0.45) / 0.225)))
        x = e.layer4(e.layer3(e.layer2(e.layer1(e.maxpool(x)))))
        x = F.relu(self.squeeze(x))
        x = self.p2(F.relu(self.p1(F.relu(self.p0(x)))))
        out = 0.01 * x.mean(3).mean(2).view(-1, 2, 1, 6)
        return out[..., :3], out[..., 3:]


def rot_from_axisangle(vec):
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = axis[..., 0:1], axis[..., 1:2], axis[..., 2:3]
    rot = torch.zeros(vec.shape[0], 4, 4, device=vec.device)
    rot[:, 0, 0], rot[:, 0, 1], rot[:, 0, 2] = (x * x * C + ca).squeeze(), (x * y * C - z * sa).squeeze(), (x * z * C + y * sa).squeeze()
    rot[:, 1, 0], rot[:, 1, 1], rot[:, 1, 2] = (x * y * C + z * sa).squeeze(), (y * y * C + ca).squeeze(), (y * z * C - x * sa).squeeze()
    rot[:, 2, 0], rot[:, 2, 1], rot[:, 2, 2] = (x * z * C - y * sa).squeeze(), (y * z * C + x * sa).squeeze(), (z * z * C + ca).squeeze()
    rot[:, 3, 3] = 1
    return rot


def transformation(axisangle, translation, invert):
    R, t = rot_from_axisangle(axisangle), translation.clone()
    if invert:
        R, t = R.transpose(1, 2), t * -1
    T = torch.zeros(t.shape[0], 4, 4, device=t.device)
    T[:, 0, 0] = T[:, 1, 1] = T[:, 2, 2] = T[:, 3, 3] = 1
    T[:, :3, 3] = t.contiguous().view(-1, 3)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


class Loss(nn.Module):
    def __init__(self, B, H, W):
        super().__init__()
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W)], 0)[None].repeat(B, 1, 1)
        self.register_buffer("pix", pix)
        self.register_buffer("ones", torch.ones(B, 1, H * W))
        self.B, self.H, self.W = B, H, W
        self.pool, self.refl = nn.AvgPool2d(3, 1), nn.ReflectionPad2d(1)

    def ssim(self, x, y):
        x, y = self.refl(x), self.refl(y)
        mx, my = self.pool(x), self.pool(y)
        sx, sy, sxy = self.pool(x * x) - mx * mx, self.pool(y * y) - my * my, self.pool(x * y) - mx * my
        n = (2 * mx * my + 1e-4) * (2 * sxy + 9e-4)
        d = (mx * mx + my * my + 1e-4) * (sx + sy + 9e-4)
        return torch.clamp((1 - n / d) / 2, 0, 1)

    def rep(self, pred, tgt):
        return 0.85 * self.ssim(pred, tgt).mean(1, True) + 0.15 * (tgt - pred).abs().mean(1, True)

    def warp(self, src, depth, K, inv_K, T):
        cam = torch.cat([depth.view(self.B, 1, -1) * torch.matmul(inv_K[:, :3, :3], self.pix), self.ones], 1)
        p = torch.matmul(torch.matmul(K, T)[:, :3, :], cam)
        xy = (p[:, :2] / (p[:, 2:3] + 1e-7)).view(self.B, 2, self.H, self.W).permute(0, 2, 3, 1).clone()
        xy[..., 0] /= self.W - 1
        xy[..., 1] /= self.H - 1
        return F.grid_sample(src, (xy - 0.5) * 2, padding_mode="border", align_corners=True)

    def forward(self, disp, tgt, srcs, Ts, K, inv_K):
        depth = 1 / (0.01 + 9.99 * disp)
        reproj = torch.cat([self.rep(self.warp(s, depth, K, inv_K, T), tgt) for s, T in zip(srcs, Ts)], 1)
        ident = torch.cat([self.rep(s, tgt) for s in srcs], 1)
        ident = ident + torch.randn(ident.shape, device=ident.device) * 1e-5
        loss = torch.min(torch.cat([ident, reproj], 1), 1)[0].mean()
        nd = disp / (disp.mean(2, True).mean(3, True) + 1e-7)
        gx, gy = (nd[..., :, :-1] - nd[..., :, 1:]).abs(), (nd[..., :-1, :] - nd[..., 1:, :]).abs()
        ix = (tgt[..., :, :-1] - tgt[..., :, 1:]).abs().mean(1, True)
        iy = (tgt[..., :-1, :] - tgt[..., 1:, :]).abs().mean(1, True)
        return loss + 1e-3 * ((gx * torch.exp(-ix)).mean() + (gy * torch.exp(-iy)).mean())


class SingleFrameStep:
    """zero_grad -> 2 pose passes + depth -> loss -> backward -> clip -> AdamW on stock torch, eager."""

    def __init__(self, B, H, W, device, lr=1e-4, weight_decay=0.01, clip=5.0, channels_last=False):
        self.depth, self.pose, self.loss = DepthNet().to(device), PoseNet().to(device), Loss(B, H, W).to(device)
        if channels_last:
            self.depth, self.pose = self.depth.to(memory_format=torch.channels_last), self.pose.to(memory_format=torch.channels_last)
        self.params = list(self.depth.parameters()) + list(self.pose.parameters())
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay)
        self.clip = clip

    def __call__(self, inputs):
        self.opt.zero_grad(set_to_none=True)
        K, inv_K = inputs[("K", 0)], inputs[("inv_K", 0)]
        aa, tr = self.pose(inputs[("color_aug", -1, 0)], inputs[("color_aug", 0, 0)])
        T_n1 = transformation(aa[:, 0], tr[:, 0], True)
        aa, tr = self.pose(inputs[("color_aug", 0, 0)], inputs[("color_aug", 1, 0)])
        T_p1 = transformation(aa[:, 0], tr[:, 0], False)
        disp = self.depth(inputs[("color_aug", 0, 0)])
        loss = self.loss(disp, inputs[("color", 0, 0)], [inputs[("color", -1, 0)], inputs[("color", 1, 0)]], [T_n1, T_p1], K, inv_K)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in self.params if p.grad is not None], self.clip)
        self.opt.step()
        return loss.detach()
