"""Seeded synthetic stereo pairs (numpy only): the tower output and the two images.

There is no network for KITTI/Middlebury or trained nets, so tests and bench.py
feed the hot path with data of the right shape and statistics:
unit-norm C-channel features where the left map is the right map shifted by a
piecewise-constant ground-truth disparity plus noise, so that argmin / LR check /
interpolation / sub-pixel all see realistic structure, and two standardised
grey images with the same shift (they drive cross() and the SGM penalties).
"""
import numpy as np


def _gauss_blur(a, sigma):
    """separable Gaussian blur (reflect padding), numpy only"""
    r = int(np.ceil(3 * sigma))
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k = (k / k.sum()).astype(np.float32)
    for axis in (0, 1):
        pad = [(0, 0), (0, 0)]
        pad[axis] = (r, r)
        p = np.pad(a, pad, mode="reflect")
        a = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), axis, p).astype(np.float32)
    return a


def natural_image(rng, H, W):
    """Piecewise-smooth grey image whose gradient statistics match the KITTI sample pair of the
    reference (standardised samples/input/kittiL.png: median |dx| 0.026, 77% of |dx| below
    tau_so = 0.08, mean cross arm 4.1 of L1 = 5 at tau1 = 0.13, mean CBCA support ~52 taps; this
    generator gives ~0.028 / 90% / 4.1 / ~50).  The roughness of the image decides how much work
    cross-based aggregation does, so white noise would understate it five-fold."""
    base = _gauss_blur(rng.standard_normal((H, W), dtype=np.float32), 15.0)
    base /= base.std()
    by, bx = 16, 48
    blk = np.kron(rng.standard_normal(((H + by - 1) // by, (W + bx - 1) // bx)).astype(np.float32),
                  np.ones((by, bx), np.float32))[:H, :W]
    tex = _gauss_blur(rng.standard_normal((H, W), dtype=np.float32), 1.0)
    tex /= tex.std()
    mask = (_gauss_blur(rng.standard_normal((H, W), dtype=np.float32), 20.0) > 0.02).astype(np.float32)
    return base + 0.8 * blk + 0.3 * tex * mask + 0.008 * rng.standard_normal((H, W), dtype=np.float32)


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

    img_scene = natural_image(rng, H, ext)
    imgR = img_scene[:, D:D + W].copy()
    imgL = img_scene[rows, cols - gt] + 0.004 * rng.standard_normal((H, W), dtype=np.float32)

    def standardise(x):  # main.lua:1095-1096 (unbiased std, as torch's :std())
        return ((x - x.mean()) / x.std(ddof=1)).astype(np.float32)

    return dict(featL=unit(featL), featR=unit(featR), imgL=standardise(imgL), imgR=standardise(imgR),
                gt=gt.astype(np.float32))
