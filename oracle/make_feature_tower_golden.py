"""Golden vectors for the feature tower (tests/golden/feature_tower.npz) -- TEST INFRASTRUCTURE.

The reference's tower is a stack of cudnn.SpatialConvolution / cudnn.ReLU modules plus nn.Normalize2 (main.lua:726-749,
Normalize2.lua:8-13); cuDNN / Torch7 cannot run here, so the generator restates the modules in PyTorch (CPU, float32):
torch.nn.functional.conv2d (cross-correlation, stride 1, zero padding 1 = the padW / padH set at main.lua:741-742), relu
between the layers, and Normalize_forward's two kernels (adcensus.cu:1284-1308) as tensor ops.  oracle/feature_tower.py
(float64 accumulation) is pinned against these vectors, the CUDA kernels against the oracle.

    python oracle/make_feature_tower_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import feature_tower as ft  # noqa: E402


def main():
    rng = np.random.default_rng(5)
    l1, fm, H, W = 4, 64, 7, 140
    layers = ft.make_weights(rng, l1=l1, fm=fm)
    x = rng.standard_normal((2, 1, H, W)).astype(np.float32)
    h = torch.from_numpy(x)
    with torch.no_grad():
        for i, (w, b) in enumerate(layers):
            h = torch.nn.functional.conv2d(h, torch.from_numpy(w), torch.from_numpy(b), stride=1, padding=1)   # main.lua:729, 741-742
            if i + 1 < l1:
                h = torch.relu(h)                                                                           # :730-732
        norm = (h * h).sum(1, keepdim=True) + 1e-5                                                         # adcensus.cu:1291-1296
        out = h / torch.sqrt(norm)                                                                         # :1305
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "feature_tower.npz")
    arrs = dict(x=x, out=out.numpy())
    for i, (w, b) in enumerate(layers):
        arrs["w%d" % i] = w
        arrs["b%d" % i] = b
    np.savez_compressed(dst, **arrs)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
