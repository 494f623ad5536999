"""Deterministic, name-keyed parameter fill shared by the golden generator (reference modules) and the tests (drop-in
modules): identical keys -> identical weights, so outputs can be compared without shipping checkpoints."""
import zlib

import torch


def fill_(module, scale=1.0):
    with torch.no_grad():
        for name, t in sorted(module.state_dict().items()):
            if not t.dtype.is_floating_point:
                continue  # num_batches_tracked
            g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
            v = torch.randn(t.shape, generator=g)
            leaf = name.rsplit(".", 1)[-1]
            if leaf == "running_var":
                v = 0.5 + torch.rand(t.shape, generator=g)
            elif leaf == "running_mean":
                v = 0.1 * v
            elif leaf in ("gamma", "gamma_xca"):
                v = 0.5 + 0.1 * v          # layer scales of order 1 so that every branch matters in the comparison
            elif leaf == "temperature":
                v = 1.0 + 0.1 * v
            elif t.dim() == 1 and leaf == "weight":
                v = 1.0 + 0.1 * v          # norm / PReLU-like vectors
            elif t.dim() == 1:
                v = 0.05 * v               # biases
            else:
                fan_in = max(1, t[0].numel())
                v = v * (scale / fan_in ** 0.5)
            t.copy_(v)
    return module


def seeded_input(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g)
