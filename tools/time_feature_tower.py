"""Time the in-library feature tower (csrc/feature_tower.cu) on a KITTI-sized stereo pair: kitti fast net shape (4 layers, 64
planes), random weights.  Run from the repository root."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mccnn_b200  # noqa: E402,F401
from mccnn_b200 import feature_tower  # noqa: E402

rng = np.random.default_rng(0)
layers, c = [], 1
for _ in range(4):
    layers.append(((rng.standard_normal((64, c, 3, 3)) / np.sqrt(9 * c)).astype(np.float32), (0.1 * rng.standard_normal(64)).astype(np.float32)))
    c = 64
t = feature_tower.FeatureTower(layers, arch="fast")
x = torch.randn((2, 1, 370, 1226), device="cuda:0")
for _ in range(3):
    f = t.forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
f = t.forward(x)
e1.record()
torch.cuda.synchronize()
flop = 2 * 2 * 370 * 1226 * (9 * 64 + 3 * 9 * 64 * 64)
ms = e0.elapsed_time(e1)
print("feature tower kitti fast (4 layers, 64 planes), 2 x 370x1226: %.3f ms  %.1f TFLOP/s useful (x3 issued on the tensor cores)" % (ms, flop / ms / 1e9))
