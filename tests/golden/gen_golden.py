"""Generate tests/golden/f1_*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:  python tests/golden/gen_golden.py
Inputs come from tests/synth.py (seeded numpy), so the tests regenerate them instead of storing them.
What is stored is what the reference computed:
  * P, T                     layers.py:28-103 transformation_from_parameters, layers.py:212 K@T
  * grid                     layers.py:192-222 (BackprojectDepth -> Project3D)
  * x0,y0                    floor of ATen's un-normalise + border clip of that grid (train.py:966-969);
                             the generator verifies that bilinear sampling from these indices reproduces
                             F.grid_sample's own output, so they are the indices ATen used
  * warp, ssim, reprojection maps, to_optimise, idx, loss   train.py:973-1051 (torch.randn patched to
                             return the seeded noise so train.py:1023 is deterministic)
  * grad_disp, grad_T, grad_axisangle, grad_translation     autograd through the reference graph
  * smooth / SI-log scalars  layers.py:231-242, train.py:924-941
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import ref_harness  # noqa: E402
import synth  # noqa: E402

import torch  # noqa: E402

torch.set_num_threads(8)

CASES = {
    # name: (seed, B, H, W, structured, with_mask, flags(no_ssim, avg, disable_automask), full)
    "tiny": (11, 2, 16, 24, False, False, (0, 0, 0), True),
    "small": (12, 2, 64, 96, True, True, (0, 0, 0), True),
    "odd": (13, 1, 21, 35, True, False, (0, 0, 0), True),
    "nossim": (14, 2, 24, 40, True, False, (1, 0, 0), True),
    "avg": (15, 2, 24, 40, True, False, (0, 1, 0), True),
    "noauto": (16, 2, 24, 40, True, True, (0, 0, 1), True),
    "avg_noauto": (17, 1, 24, 40, True, False, (0, 1, 1), True),
    "cfg1": (1234, 2, 128, 416, True, False, (0, 0, 0), False),
    # BASELINE.json configs 2-5 at their full sizes (sub-sampled maps, stride 499; full argmin map)
    "cfg2": (1235, 12, 192, 640, True, False, (0, 0, 0), 499),
    "cfg4": (1236, 6, 320, 1024, True, True, (0, 0, 0), 499),
    "cfg5": (1237, 8, 384, 1280, True, False, (0, 0, 0), 499),
}


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def run_case(T, name, spec):
    seed, B, H, W, structured, with_mask, fl, full = spec
    nid = 1 if fl[1] else 2
    c = synth.make_case(seed, B, H, W, structured, with_mask, nid)
    tr = ref_harness.make_trainer(T, B, H, W)
    tr.opt.no_ssim, tr.opt.avg_reprojection, tr.opt.disable_automasking = [bool(x) for x in fl]
    tt = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else None) for k, v in c.items()}
    disp = tt["disp"].clone().requires_grad_(True)
    aa = tt["axisangle"].clone().requires_grad_(True)
    tl = tt["translation"].clone().requires_grad_(True)
    K, inv_K = tt["K"], tt["inv_K"]
    out = {}
    Ts, Ps, grids, warps = [], [], [], []
    srcs = [tt["src0"], tt["src1"]]
    for k in range(2):
        # source 0 uses the forward transform, source 1 the inverted one (both code paths of layers.py:28-45)
        Tm = T.transformation_from_parameters(aa[k], tl[k], invert=(k == 1))
        Tm.retain_grad()
        Ts.append(Tm)
        Ps.append(torch.matmul(K, Tm)[:, :3, :].detach())
        # generate_images_pred body, split so the grid can be captured (train.py:960-969)
        _, depth = T.disp_to_depth(disp, tr.opt.min_depth, tr.opt.max_depth)
        cam = tr.backproject_depth(depth, inv_K)
        pix = tr.project_3d(cam, K, Tm)
        grids.append(pix.detach().clone())
        w = T.F.grid_sample(srcs[k], pix, padding_mode="border", align_corners=True)
        # the unsplit reference method must agree bit for bit
        w2 = tr.generate_images_pred({("disp", 0): disp}, Tm, srcs[k], K, inv_K)
        assert torch.equal(w, w2)
        warps.append(w)
    noise = tt["noise"]
    real_randn = torch.randn
    torch.randn = lambda *a, **kw: noise.clone()
    try:
        loss, auto_mask = tr.compute_losses_base({("disp", 0): disp}, tt["tgt"], warps, srcs, tt["mask_rec"])
    finally:
        torch.randn = real_randn
    loss.backward()
    # indices ATen used (validated against its output below)
    xy = np.zeros((2, 2, B, H, W), np.int32)
    for k in range(2):
        g = grids[k]
        ix = ((g[..., 0] + 1) * ((W - 1) / 2)).clamp(0, W - 1)
        iy = ((g[..., 1] + 1) * ((H - 1) / 2)).clamp(0, H - 1)
        x0, y0 = ix.floor(), iy.floor()
        xy[k, 0], xy[k, 1] = x0.numpy().astype(np.int32), y0.numpy().astype(np.int32)
        fw, fn = ix - x0, iy - y0
        x0l, y0l = x0.long(), y0.long()
        x1l, y1l = (x0l + 1).clamp(max=W - 1), (y0l + 1).clamp(max=H - 1)
        src = srcs[k]
        bi = torch.arange(B).view(B, 1, 1)

        def gat(yy, xx):
            return src[bi, :, yy, xx].permute(0, 3, 1, 2)
        e, s = 1 - fw, 1 - fn
        rec = gat(y0l, x0l) * (s * e).unsqueeze(1) + gat(y0l, x1l) * (s * fw).unsqueeze(1) * ((x0l + 1) < W).unsqueeze(1) \
            + gat(y1l, x0l) * (fn * e).unsqueeze(1) * ((y0l + 1) < H).unsqueeze(1) \
            + gat(y1l, x1l) * (fn * fw).unsqueeze(1) * (((x0l + 1) < W) & ((y0l + 1) < H)).unsqueeze(1)
        err = (rec - warps[k].detach()).abs().max().item()
        assert err < 2e-6, ("index validation failed", name, err)
    # per-pixel maps (reference methods; combination restated from train.py:1023-1036 with the same noise)
    with torch.no_grad():
        rep = [tr.compute_reprojection_loss(w.detach(), tt["tgt"]) for w in warps]
        idl = [tr.compute_reprojection_loss(s, tt["tgt"]) for s in srcs]
        rep_c, idl_c = torch.cat(rep, 1), torch.cat(idl, 1)
        if fl[1]:
            rep_c, idl_c = rep_c.mean(1, keepdim=True), idl_c.mean(1, keepdim=True)
        comb = rep_c if fl[2] else torch.cat((idl_c + noise * 0.00001, rep_c), 1)
        if comb.shape[1] == 1:
            to_opt, idx = comb[:, 0], torch.zeros_like(comb[:, 0], dtype=torch.long)
        else:
            to_opt, idx = torch.min(comb, dim=1)
        if tt["mask_rec"] is not None:
            to_opt = to_opt * tt["mask_rec"][:, 0]
        mean_disp = disp.mean(2, True).mean(3, True)
        smooth = T.get_smooth_loss(disp / (mean_disp + 1e-7), tt["tgt"])
        photo = to_opt.mean()
        assert abs((photo + tr.opt.disparity_smoothness * smooth).item() - loss.item()) < 1e-6
        ssim0 = tr.ssim(warps[0].detach(), tt["tgt"])
        _, depth = T.disp_to_depth(disp, tr.opt.min_depth, tr.opt.max_depth)
        depth2 = depth * (1.0 + 0.1 * torch.from_numpy(c["noise"][:, :1]).clamp(-3, 3) * 0.3)
        si = tr.compute_SI_log_depth_loss(depth, depth2)
        si_m = tr.compute_SI_log_depth_loss(depth, depth2, (tt["noise"][:, :1] > 0).float())
    out.update(
        spec=np.array([seed, B, H, W, int(structured), int(with_mask), fl[0], fl[1], fl[2]], np.int64),
        loss=np.float64(loss.item()), photo=np.float64(photo.item()), smooth=np.float64(smooth.item()),
        si_log=np.float64(si.item()), si_log_masked=np.float64(si_m.item()),
        T=torch.stack([t.detach() for t in Ts]).numpy(), P=torch.stack(Ps).numpy(),
        grad_T=torch.stack([t.grad for t in Ts]).numpy(),
        grad_axisangle=aa.grad.numpy(), grad_translation=tl.grad.numpy(),
        x0y0_crc=crc(xy), x0y0_sum=np.int64(xy.astype(np.int64).sum()),
        grid_crc=np.array([crc(g.numpy()) for g in grids], np.uint32),
        idx=idx.numpy().astype(np.uint8),
        grad_disp_abs_sum=np.float64(disp.grad.abs().double().sum().item()),
    )
    if full is True:
        out.update(grid=torch.stack(grids).numpy(), x0y0=xy, warp=torch.stack([w.detach() for w in warps]).numpy(),
                   ssim0=ssim0.numpy(), rep=torch.cat(rep, 1).numpy(), idl=torch.cat(idl, 1).numpy(),
                   to_optimise=to_opt.numpy(), grad_disp=disp.grad.numpy(),
                   depth=depth.detach().numpy())
    else:
        st = 37 if full is False else int(full)
        out.update(stride=np.int64(st), to_optimise_sub=to_opt.numpy().ravel()[::st].copy(),
                   grad_disp_sub=disp.grad.numpy().ravel()[::st].copy(),
                   warp_sub=torch.stack([w.detach() for w in warps]).numpy().ravel()[::st].copy(),
                   grid_sub=torch.stack(grids).numpy().ravel()[::st].copy(),
                   x0y0_sub=xy.ravel()[::st].copy())
    path = os.path.join(HERE, "f1_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-10s loss=%.6f photo=%.6f smooth=%.6f idx_hist=%s  -> %s (%.0f KB)" % (
        name, out["loss"], out["photo"], out["smooth"], np.bincount(out["idx"].ravel(), minlength=4),
        os.path.basename(path), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    only = sys.argv[1:]
    T = ref_harness.import_reference(192, 640, 2)
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        run_case(T, name, spec)
