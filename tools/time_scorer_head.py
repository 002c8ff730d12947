"""Time the fused scorer head (csrc/scorer_head.cu) at the KITTI-accurate size: ms per call, TFLOP/s of the useful MLP flops
(one pass) and of the issued tensor-core flops (x3 for the bf16 split)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mccnn_b200  # noqa: E402,F401
from mccnn_b200 import scorer_head  # noqa: E402
from oracle import scorer_head as osh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--H", type=int, default=370)
ap.add_argument("--W", type=int, default=1226)
ap.add_argument("--D", type=int, default=228)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--l2", type=int, default=4)
a = ap.parse_args()
fm, nh2 = 112, 384
rng = np.random.default_rng(0)
layers = osh.make_weights(rng, fm, nh2, a.l2)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
fL = torch.relu(torch.randn((fm, a.H, a.W), device=dev, generator=g))
fR = torch.relu(torch.randn((fm, a.H, a.W), device=dev, generator=g))
head = scorer_head.ScorerHead(layers)
rows = a.H * (a.D * a.W - a.D * (a.D - 1) // 2)
flop_row = 2 * (2 * fm * nh2 + (a.l2 - 1) * nh2 * nh2 + nh2)
for nterms in (3, 1):
    head.volumes(fL, fR, a.D, nterms=nterms)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        head.volumes(fL, fR, a.D, nterms=nterms)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    print("scorer head %dx%d d=%d l2=%d nterms=%d: %.2f ms  useful %.1f TFLOP/s  issued %.1f TFLOP/s (bf16 tensor)"
          % (a.H, a.W, a.D, a.l2, nterms, ms, rows * flop_row / ms / 1e9, rows * flop_row * nterms / ms / 1e9))
head.close()
