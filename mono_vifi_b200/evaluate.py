"""Evaluation loop of the reference's trainer (train.py:419-483 `test_kitti`; the same arithmetic as evaluate_depth.py:134-193) on
the GPU: inference through the drop-in encoder / decoder (tcgen05 convolutions, inference BatchNorm kernel), then per image ONE call
of mvf_depth_eval -- bilinear resize of the predicted disparity to the ground-truth size, inversion, validity mask (Eigen crop),
median scaling, clamp and the seven metrics of compute_depth_errors (layers.py:293-311) -- with no host synchronisation until the
final averages.  The reference walks the images in Python with ~25 small launches and a boolean-mask compaction each."""
import numpy as np
import torch

from . import _lib
from . import layers as L

MIN_DEPTH, MAX_DEPTH, STEREO_SCALE_FACTOR = 1e-3, 80.0, 5.4     # train.py:425-428
METRICS = ("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3")


def depth_metrics(pred_disp, gt_depth, eval_split="eigen", use_stereo=False, out=None, workspace=None):
    """pred_disp [h,w] (scaled disparity), gt_depth [Hg,Wg] (both CUDA fp32) -> [8] = the seven metrics + the scale ratio used."""
    if not pred_disp.is_cuda:
        raise RuntimeError("depth_metrics: CUDA tensors only (there is no CPU fallback)")
    lib = _lib.lib()
    pred_disp, gt_depth = pred_disp.contiguous().float(), gt_depth.contiguous().float()
    h, w = pred_disp.shape[-2:]
    Hg, Wg = gt_depth.shape[-2:]
    n = lib.mvf_depth_eval_workspace_bytes(Hg, Wg)
    if workspace is None or workspace.numel() < n:
        workspace = torch.empty(n, device=pred_disp.device, dtype=torch.uint8)
    if out is None:
        out = torch.empty(8, device=pred_disp.device, dtype=torch.float32)
    _lib.check(lib.mvf_depth_eval(pred_disp.data_ptr(), h, w, gt_depth.data_ptr(), Hg, Wg, MIN_DEPTH, MAX_DEPTH,
                                  1 if eval_split == "eigen" else 0, STEREO_SCALE_FACTOR if use_stereo else 0.0, workspace.data_ptr(),
                                  workspace.numel(), out.data_ptr(), torch.cuda.current_stream(pred_disp.device).cuda_stream), "mvf_depth_eval")
    return out


@torch.no_grad()
def evaluate_depth(models, batches, gt_depths, opt, eval_split="eigen", use_stereo=False):
    """models: {"encoder", "depth"} (the drop-in networks); batches: iterable of dicts holding ("color", 0, 0) [B,3,H,W];
    gt_depths: sequence of [Hg,Wg] arrays / tensors, one per image in order.  Returns the mean metrics and the scale statistics,
    as train.py:470-481 logs them."""
    was_training = {k: m.training for k, m in models.items()}
    for m in models.values():
        m.eval()
    dev = next(models["encoder"].parameters()).device
    n_img = len(gt_depths)
    results = torch.empty(n_img, 8, device=dev, dtype=torch.float32)
    workspace = None
    i = 0
    for data in batches:
        color = data[("color", 0, 0)].to(dev, non_blocking=True)
        disp = models["depth"](models["encoder"](color))[("disp", 0)]
        pred_disp, _ = L.disp_to_depth(disp, opt.min_depth, opt.max_depth)
        for b in range(pred_disp.shape[0]):
            if i >= n_img:
                break
            gt = gt_depths[i]
            gt = (torch.from_numpy(np.ascontiguousarray(gt)) if isinstance(gt, np.ndarray) else gt).to(dev, non_blocking=True).float()
            need = _lib.lib().mvf_depth_eval_workspace_bytes(gt.shape[0], gt.shape[1])
            if workspace is None or workspace.numel() < need:
                workspace = torch.empty(need, device=dev, dtype=torch.uint8)
            depth_metrics(pred_disp[b, 0], gt, eval_split, use_stereo, out=results[i], workspace=workspace)
            i += 1
    for k, m in models.items():
        m.train(was_training[k])
    res = results[:i].cpu()                       # the only synchronisation
    out = {k: float(res[:, j].mean()) for j, k in enumerate(METRICS)}
    if not use_stereo:
        ratios = res[:, 7]
        med = torch.median(ratios)
        out["scale_med"], out["scale_std"] = float(med), float(torch.std(ratios / med)) if i > 1 else 0.0
    out["n_images"] = i
    return out
