"""The tensor transforms of the affine-augmentation branch of `process_batch` (train.py:815-922),
batched and free of host synchronisation.

The reference walks the batch in Python and reads `angle` / `box` with `.item()` (one device->host sync per value,
train.py:889-897, 906-913), which also rules out recording the step into a CUDA graph.  Here the same arithmetic runs on
whole batches with the per-sample angle / box kept as device tensors:

* `rotate`            = torchvision.transforms.functional.rotate(img, angle, interpolation=BILINEAR) per sample
                        (inverse rotation about the image centre -> normalised grid -> grid_sample, zeros outside);
* `crop_resize`       = img[:, :, y0:y0+h, x0:x0+w] followed by F.interpolate(.., [H, W], "bilinear",
                        align_corners=False): source index max(scale * (dst + 0.5) - 0.5, 0), right / lower neighbour
                        clamped to the CROP (not to the image);
* `affine_transform`  = Trainer.affine_transform (train.py:887-901): rotate -> crop -> resize;
* `depth_restore`     = the first half of Trainer.compute_depth_consistency_loss_affine (train.py:903-916): shrink the
                        affine depth to the box, paste it into a zero canvas, rotate back by -angle, scale by ratio_local;
* `conjugate_pose`    = Rt_Rc of train.py:820-829: [Rc R Rc^-1 | Rc t] with a zero last row.

Everything is plain torch (autograd provides the adjoints); the fused CUDA kernels of the loss groups and of the SI-log
term consume the results.  Parity with the reference's own methods: tests/test_affine.py (tests/golden/affine_golden.npz).
"""
import math

import torch
import torch.nn.functional as F


def _theta(angle_deg):
    """[B] degrees -> [B,2,3] float32 inverse-rotation matrix torchvision builds for rotate(angle): centre 0,
    _get_inverse_affine_matrix(.., -angle, ..) = [cos r, sin r, 0, -sin r, cos r, 0] with r = radians(-angle) in double"""
    r = (-angle_deg.reshape(-1).double()) * (math.pi / 180.0)
    c, s = torch.cos(r), torch.sin(r)
    z = torch.zeros_like(c)
    return torch.stack([torch.stack([c, s, z], 1), torch.stack([-s, c, z], 1)], 1).float()


def rotate(img, angle_deg):
    """img [B,C,H,W], angle_deg [B] or [B,1] (counter-clockwise, degrees) -> rotated about the centre, bilinear, zeros"""
    B, _, H, W = img.shape
    dev = img.device
    theta = _theta(angle_deg).to(dev)
    base = torch.empty(1, H, W, 3, dtype=torch.float32, device=dev)
    base[..., 0].copy_(torch.linspace(-W * 0.5 + 0.5, W * 0.5 + 0.5 - 1, steps=W, device=dev))
    base[..., 1].copy_(torch.linspace(-H * 0.5 + 0.5, H * 0.5 + 0.5 - 1, steps=H, device=dev).unsqueeze(-1))
    base[..., 2].fill_(1)
    # (fill kernels, not a host->device copy: the step may be inside a CUDA-graph capture)
    half = torch.stack([torch.full((), 0.5 * W, dtype=torch.float32, device=dev), torch.full((), 0.5 * H, dtype=torch.float32, device=dev)])
    rescaled = theta.transpose(1, 2) / half
    grid = base.view(1, H * W, 3).expand(B, H * W, 3).bmm(rescaled).view(B, H, W, 2)
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


def _linear_taps(n_out, n_in, device):
    """ATen's bilinear (align_corners=False) source taps for resizing n_in[b] -> n_out samples.
    n_in: [B] float32.  Returns (i0, i1 as int64 [B, n_out], lambda1 [B, n_out])."""
    scale = n_in / float(n_out)                                            # float32, as area_pixel_compute_scale<float>
    dst = torch.arange(n_out, device=device, dtype=torch.float32)
    src = (scale[:, None] * (dst[None, :] + 0.5) - 0.5).clamp_min(0.0)
    i0 = src.floor()
    lam = src - i0
    i0 = i0.long()
    last = (n_in.long() - 1)[:, None]
    i0 = torch.minimum(i0, last)
    i1 = i0 + (i0 < last).long()
    return i0, i1, lam


def _gather2d(img, iy, ix):
    """img [B,C,H,W], iy [B,Ho], ix [B,Wo] (int64) -> [B,C,Ho,Wo] = img[b, :, iy[b,i], ix[b,j]]"""
    B, C, H, W = img.shape
    rows = torch.gather(img, 2, iy[:, None, :, None].expand(B, C, iy.shape[1], W))
    return torch.gather(rows, 3, ix[:, None, None, :].expand(B, C, iy.shape[1], ix.shape[1]))


def crop_resize(img, box, out_hw):
    """F.interpolate(img[b, :, y0:y0+h, x0:x0+w], out_hw, 'bilinear', align_corners=False) for every sample;
    box [B,4] = (x0, y0, w, h) integers on the device"""
    Ho, Wo = out_hw
    box = box.to(img.device)
    x0, y0, w, h = box[:, 0].long(), box[:, 1].long(), box[:, 2].float(), box[:, 3].float()
    ix0, ix1, lx = _linear_taps(Wo, w, img.device)
    iy0, iy1, ly = _linear_taps(Ho, h, img.device)
    ix0, ix1, iy0, iy1 = ix0 + x0[:, None], ix1 + x0[:, None], iy0 + y0[:, None], iy1 + y0[:, None]
    lx, ly = lx[:, None, None, :], ly[:, None, :, None]
    top = _gather2d(img, iy0, ix0) * (1 - lx) + _gather2d(img, iy0, ix1) * lx
    bot = _gather2d(img, iy1, ix0) * (1 - lx) + _gather2d(img, iy1, ix1) * lx
    return top * (1 - ly) + bot * ly


def affine_transform(img, angle, box):
    """Trainer.affine_transform (train.py:887-901): rotate by angle, crop the box, resize back to the input size"""
    return crop_resize(rotate(img, angle), box, img.shape[-2:])


def shrink_paste(x, box):
    """canvas[b, :, y0:y0+h, x0:x0+w] = F.interpolate(x[b], [h, w], 'bilinear', align_corners=False), zeros elsewhere
    (train.py:911-913); the canvas has the size of x"""
    B, C, H, W = x.shape
    dev = x.device
    box = box.to(dev)
    x0, y0, w, h = box[:, 0].long(), box[:, 1].long(), box[:, 2].long(), box[:, 3].long()

    def taps(n_canvas, n_src, o, n_box):
        # resize n_src -> n_box[b] samples, evaluated at canvas position p (destination index p - o[b])
        scale = float(n_src) / n_box.float()
        p = torch.arange(n_canvas, device=dev)
        d = p[None, :] - o[:, None]
        inside = (d >= 0) & (d < n_box[:, None])
        src = (scale[:, None] * (d.float() + 0.5) - 0.5).clamp_min(0.0)
        i0 = src.floor()
        lam = src - i0
        i0 = i0.long().clamp(0, n_src - 1)
        i1 = i0 + (i0 < n_src - 1).long()
        return i0, i1, lam, inside

    ix0, ix1, lx, inx = taps(W, W, x0, w)
    iy0, iy1, ly, iny = taps(H, H, y0, h)
    lx, ly = lx[:, None, None, :], ly[:, None, :, None]
    top = _gather2d(x, iy0, ix0) * (1 - lx) + _gather2d(x, iy0, ix1) * lx
    bot = _gather2d(x, iy1, ix0) * (1 - lx) + _gather2d(x, iy1, ix1) * lx
    inside = (iny[:, None, :, None] & inx[:, None, None, :]).to(x.dtype)
    return (top * (1 - ly) + bot * ly) * inside


def depth_restore(depth_affine, angle, box, ratio_local):
    """depth of the augmented view mapped back onto the original view (train.py:911-916)"""
    canvas = shrink_paste(depth_affine, box)
    back = rotate(canvas, -angle.reshape(-1))
    return back * ratio_local.reshape(-1, 1, 1, 1).to(back.dtype)


def conjugate_pose(pose, Rc):
    """Rt_Rc (train.py:820-829): [:3,:3] = Rc R Rc^-1, [:3,3] = Rc t, everything else (the last row too) zero"""
    out = torch.zeros_like(pose)
    out[:, :3, :3] = torch.matmul(Rc, torch.matmul(pose[:, :3, :3], torch.linalg.inv_ex(Rc)[0]))
    out[:, :3, 3:4] = torch.matmul(Rc, pose[:, :3, 3:4])
    return out


def augmentation_geometry(K, inv_K, angle, ratio, origin, H, W):
    """What the data loader derives for one affine augmentation (datasets/mono_dataset.py:110-136), for a whole batch:
    the image is enlarged by `ratio` (> 1), rotated by `angle` degrees about its centre and an H x W window is cut at
    `origin` (fractions in [0,1) of the slack in x and y).  Returns
      Rc   [B,3,3]  camera-space matrix with  p_aug ~ K Rc K^-1 p  (rotation R in the image plane conjugated by K, plus the
                    crop / scale offset folded into the third column),
      box  [B,4]    the window (x0, y0, w, h) in ORIGINAL-resolution pixels, i.e. divided by ratio and rounded.
    K / inv_K [B,4,4] (or [B,3,3]), angle / ratio [B] or [B,1], origin [B,2]."""
    B = angle.shape[0]
    angle, ratio = angle.reshape(B).double(), ratio.reshape(B).double()
    K3, iK3 = K[:, :3, :3].float().cpu(), inv_K[:, :3, :3].float().cpu()
    He, We = (H * ratio).floor(), (W * ratio).floor()                      # enlarged size, truncated like int()
    w0 = ((We - W) * origin[:, 0].double()).floor()
    h0 = ((He - H) * origin[:, 1].double()).floor()
    rad = angle * (math.pi / 180.0)
    c, s_ = torch.cos(rad), torch.sin(rad)
    R = torch.zeros(B, 3, 3, dtype=torch.float64)
    R[:, 0, 0], R[:, 0, 1], R[:, 1, 0], R[:, 1, 1], R[:, 2, 2] = c, s_, -s_, c, 1.0
    R = R.float()
    fs = 1.0 / ratio
    centre = torch.stack([-fs * We / 2, -fs * He / 2, fs - 1], 1).float()
    shift = torch.stack([(We / 2 - w0) * fs, (He / 2 - h0) * fs, torch.zeros(B, dtype=torch.float64)], 1).float()
    offset = torch.bmm(R, centre[:, :, None])[:, :, 0] + shift
    Rc = torch.bmm(torch.bmm(iK3, R), K3)
    Rc[:, :, 2] += torch.bmm(iK3, offset[:, :, None])[:, :, 0]
    rnd = lambda v: torch.round(v).long()                                   # Python round(): half to even, as torch.round
    box = torch.stack([rnd(w0 / ratio), rnd(h0 / ratio), rnd(W / ratio), rnd(H / ratio)], 1)
    return Rc, box

