"""Import the UNMODIFIED reference (/root/reference) on CPU so golden vectors can be generated.

Only usable in the build container (the reference is not present on the GPU box).  Nothing in
`tests/`, `bench.py` or `__graft_entry__.py` imports this at run time; it is the committed
generator side of `tests/golden/*.npz` (see gen_golden.py).

Recipe follows SURVEY.md §9.1: shim the absent pure-Python third-party modules (none of them
touches arithmetic on the hot path), import `train`, and build a `Trainer` through `__new__`
plus attribute injection so the reference's own methods (train.py:924-1051) run on CPU.
"""
import argparse
import os
import sys
import types

REF = "/root/reference"


def _install_shims():
    import torch
    import torch.nn as nn

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        # any other attribute resolves to an inert placeholder (dead imports in the reference)
        def _ga(k):
            if k.startswith("__"):
                raise AttributeError(k)
            return lambda *a, **kw: None
        m.__getattr__ = _ga
        m.__path__ = []
        sys.modules[name] = m
        return m

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1.0 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            return x * mask / keep

    if "timm" not in sys.modules:
        mod("timm")
        mod("timm.models")
        mod("timm.models.layers", DropPath=DropPath, trunc_normal_=torch.nn.init.trunc_normal_)

    class CfgNode(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

        def defrost(self):
            pass

        def freeze(self):
            pass

        def merge_from_file(self, *_):
            pass

        def merge_from_list(self, *_):
            pass

    if "yacs" not in sys.modules:
        mod("yacs")
        mod("yacs.config", CfgNode=CfgNode)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "skimage",
                 "skimage.transform", "h5py", "imageio"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod(name)
    if "thop" not in sys.modules:
        try:
            import thop  # noqa: F401
        except Exception:
            mod("thop", profile=lambda *a, **k: (0, 0), clever_format=lambda x, f: x)

    class ArgumentParser(argparse.ArgumentParser):
        def add_argument(self, *args, **kw):
            kw.pop("is_config_file", None)
            return super().add_argument(*args, **kw)

        def parse_args(self, args=None, namespace=None):
            args = list(sys.argv[1:] if args is None else args)
            out = []
            i = 0
            cfg = None
            while i < len(args):
                if args[i] in ("-c", "--config"):
                    cfg = args[i + 1]
                    i += 2
                else:
                    out.append(args[i])
                    i += 1
            pre = []
            if cfg is not None:
                pre += ["-c", cfg]
                for line in open(cfg):
                    line = line.strip()
                    if not line or line.startswith("#") or "=" not in line:
                        continue
                    k, v = [s.strip() for s in line.split("=", 1)]
                    if v == "True":
                        pre.append("--" + k)
                    elif v == "False":
                        continue
                    else:
                        pre += ["--" + k] + v.strip("[]").replace(",", " ").split()
            return super().parse_args(pre + out, namespace)

    if "configargparse" not in sys.modules:
        try:
            import configargparse  # noqa: F401
        except Exception:
            mod("configargparse", ArgumentParser=ArgumentParser)


def import_reference(height=192, width=640, batch_size=2, extra=()):
    """Returns the reference's `train` module (with `opts` parsed for the given shape)."""
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree not present (golden generation only runs in the build container)")
    _install_shims()
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    sys.argv = ["train.py", "-c", os.path.join(REF, "configs/resnet18/ResNet18_KITTI_MR.txt"),
                "--height", str(height), "--width", str(width), "--batch_size", str(batch_size),
                "--weights_init", "scratch"] + list(extra)
    import train as T  # noqa
    return T


def make_trainer(T, B, H, W, device="cpu", with_models=False, seed=1234):
    """Trainer.__new__ + attribute injection."""
    import copy
    import torch
    tr = T.Trainer.__new__(T.Trainer)
    opt = copy.copy(T.opts)
    opt.height, opt.width, opt.batch_size = H, W, B
    opt.use_affine = False
    tr.opt = opt
    tr.device = torch.device(device)
    tr.ssim = T.SSIM()
    tr.backproject_depth = T.BackprojectDepth(B, H, W)
    tr.project_3d = T.Project3D(B, H, W)
    if with_models:
        torch.manual_seed(seed)
        import networks
        tr.models = {}
        tr.models["encoder"] = networks.monodepth2.DepthEncoder(18, False)
        tr.models["depth"] = networks.monodepth2.DepthDecoder(tr.models["encoder"].num_ch_enc, range(1))
        tr.models["pose_encoder"] = networks.posenet.ResnetEncoder(18, False, num_input_images=2)
        tr.models["pose"] = networks.posenet.PoseDecoder(tr.models["pose_encoder"].num_ch_enc,
                                                         num_input_features=1, num_frames_to_predict_for=2)
    return tr
