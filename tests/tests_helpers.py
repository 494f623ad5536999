"""Test-side helpers shared by the GPU tests, tools and bench (not under test)."""
import numpy as np


def synth_T(axisangle, translation, invert):
    """layers.py:28-103 restated in float64 numpy, to make plausible [B,4,4] poses for synthetic cases."""
    B = axisangle.shape[0]
    out = np.zeros((B, 4, 4), np.float64)
    for b in range(B):
        v = axisangle[b, 0].astype(np.float64)
        t = translation[b, 0].astype(np.float64)
        ang = np.linalg.norm(v)
        ax = v / (ang + 1e-7)
        Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        R = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
        M = np.eye(4)
        if invert:
            M[:3, :3] = R.T
            M[:3, 3] = -R.T @ t
        else:
            M[:3, :3] = R
            M[:3, 3] = t
        out[b] = M
    return out.astype(np.float32)
