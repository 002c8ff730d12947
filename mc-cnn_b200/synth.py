"""Seeded synthetic stereo pairs (numpy only): the tower output and the two images.

There is no network for KITTI/Middlebury or trained nets, so tests and bench.py
feed the hot path with data of the right shape and statistics:
unit-norm C-channel features where the left map is the right map shifted by a
piecewise-constant ground-truth disparity plus noise, so that argmin / LR check /
interpolation / sub-pixel all see realistic structure, and two standardised
grey images with the same shift (they drive cross() and the SGM penalties).
"""
import numpy as np


def _box3(a):
    p = np.pad(a, ((1, 1), (1, 1)), mode="edge")
    out = np.zeros_like(a)
    for dy in range(3):
        for dx in range(3):
            out += p[dy:dy + a.shape[0], dx:dx + a.shape[1]]
    return out / 9.0


def make_pair(H, W, C, D, seed=0, noise=0.1, block=(16, 48)):
    """Returns dict(featL, featR (C,H,W), imgL, imgR (H,W), gt (H,W)) as float32 arrays."""
    rng = np.random.default_rng(seed)
    dmax = max(1, min(D, max(2, W // 3)))
    by, bx = block
    gy, gx = (H + by - 1) // by, (W + bx - 1) // bx
    gt_blocks = rng.integers(0, dmax, size=(gy, gx))
    gt = np.kron(gt_blocks, np.ones((by, bx), dtype=np.int64))[:H, :W]

    ext = W + D
    scene = rng.standard_normal((C, H, ext), dtype=np.float32)
    cols = np.arange(W)[None, :] + D
    rows = np.arange(H)[:, None]
    featR = scene[:, :, D:D + W].copy()
    featL = scene[:, rows, cols - gt] + noise * rng.standard_normal((C, H, W), dtype=np.float32)

    def unit(f):  # Normalize2 (adcensus.cu:1284-1308)
        return (f / np.sqrt((f * f).sum(0, keepdims=True) + 1e-5)).astype(np.float32)

    img_scene = rng.standard_normal((H, ext), dtype=np.float32)
    img_scene = _box3(_box3(img_scene))
    imgR = img_scene[:, D:D + W].copy()
    imgL = img_scene[rows, cols - gt] + 0.02 * rng.standard_normal((H, W), dtype=np.float32)

    def standardise(x):  # main.lua:1095-1096 (unbiased std, as torch's :std())
        return ((x - x.mean()) / x.std(ddof=1)).astype(np.float32)

    return dict(featL=unit(featL), featR=unit(featR), imgL=standardise(imgL), imgR=standardise(imgR),
                gt=gt.astype(np.float32))
