"""Seeded synthetic inputs for the view-synthesis / photometric-loss path.

Pure numpy (PCG64 streams are stable across numpy versions), so the golden generator (build
container, reference importable) and the tests (anywhere) regenerate identical arrays from a seed.
"""
import numpy as np


def kitti_K(B, H, W):
    """kitti_dataset.py:23-26 normalised intrinsics scaled to (W,H); mono_dataset.py:243-252."""
    K = np.array([[0.58 * W, 0, 0.5 * W, 0],
                  [0, 1.92 * H, 0.5 * H, 0],
                  [0, 0, 1, 0],
                  [0, 0, 0, 1]], dtype=np.float32)
    inv_K = np.linalg.pinv(K)
    return np.repeat(K[None], B, 0).copy(), np.repeat(inv_K[None], B, 0).astype(np.float32).copy()


def _texture(rng, B, H, W, shift):
    """Low-frequency sinusoid texture + mild noise, translated by `shift` px (argmin not noise-dominated)."""
    v, u = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    out = np.zeros((B, 3, H, W), np.float64)
    for b in range(B):
        for c in range(3):
            acc = np.zeros((H, W))
            for _ in range(4):
                fx, fy = rng.uniform(0.02, 0.25, 2)
                ph = rng.uniform(0, 2 * np.pi)
                acc += rng.uniform(0.3, 1.0) * np.sin(fx * (u + shift[b, 0]) + fy * (v + shift[b, 1]) + ph)
            out[b, c] = 0.5 + 0.11 * acc
    return out


def make_case(seed, B, H, W, structured=True, with_mask=False, nid=2):
    """Returns dict of fp32 arrays: disp, tgt, src0, src1, K, inv_K, axisangle[2], translation[2], noise, mask_rec."""
    rng = np.random.default_rng(seed)
    K, inv_K = kitti_K(B, H, W)
    if structured:
        trng = np.random.default_rng(seed + 7919)
        state = trng.bit_generator.state
        imgs = []
        for s in (0.0, -2.0, 2.0):
            trng.bit_generator.state = state  # same texture, shifted
            sh = np.tile(np.array([[s, 0.5 * s]]), (B, 1))
            imgs.append(_texture(trng, B, H, W, sh))
        tgt, src0, src1 = [np.clip(im + 0.02 * rng.standard_normal(im.shape), 0, 1).astype(np.float32)
                           for im in imgs]
        v, u = np.meshgrid(np.linspace(0, 1, H), np.linspace(0, 1, W), indexing="ij")
        disp = 0.15 + 0.6 * v[None, None] * np.ones((B, 1, 1, 1)) + 0.1 * np.sin(6 * u)[None, None]
        disp = np.clip(disp + 0.03 * rng.standard_normal(disp.shape), 0.01, 0.99).astype(np.float32)
    else:
        tgt = rng.random((B, 3, H, W), dtype=np.float32)
        src0 = rng.random((B, 3, H, W), dtype=np.float32)
        src1 = rng.random((B, 3, H, W), dtype=np.float32)
        disp = rng.random((B, 1, H, W), dtype=np.float32)
    axisangle = (0.01 * rng.standard_normal((2, B, 1, 3))).astype(np.float32)
    translation = (0.05 * rng.standard_normal((2, B, 1, 3))).astype(np.float32)
    noise = rng.standard_normal((B, nid, H, W)).astype(np.float32)
    mask_rec = None
    if with_mask:
        mask_rec = (rng.random((B, 1, H, W)) > 0.2).astype(np.float32)
    return dict(disp=disp, tgt=tgt, src0=src0, src1=src1, K=K, inv_K=inv_K, axisangle=axisangle,
                translation=translation, noise=noise, mask_rec=mask_rec)
