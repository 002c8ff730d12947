"""Golden vectors for the accurate architecture's scorer head (tests/golden/scorer_head.npz) -- TEST INFRASTRUCTURE.

The reference's head is a Lua module chain on cuBLAS / cuDNN (main.lua:688-695, 958-984; SpatialConvolution1_fw.lua:11-31);
Torch7 cannot run here, so the generator restates those lines operation by operation in PyTorch (CPU, float32 -- the
reference's dtype): per disparity the slices / copies of main.lua:963-975, `output[i]:addmm(0, 1, weight, input[i])` +
`output:add(bias:expandAs(output))` per layer, cudnn.ReLU / cudnn.Sigmoid.  oracle/scorer_head.py (float64 accumulation) is
pinned against these vectors by tests/test_scorer_head_oracle.py, and the CUDA kernel against the oracle.

    python oracle/make_scorer_head_golden.py        # rewrites tests/golden/scorer_head.npz
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import scorer_head as sh  # noqa: E402


def conv1_fw(x, weight, bias):
    """SpatialConvolution1_fw.lua:11-31 for num_ex = 1: x (1, fm_in, h, w) -> (1, fm_out, h, w)"""
    _, fm_in, h, w = x.shape
    out = torch.zeros((weight.shape[0], h * w), dtype=torch.float32)
    out = torch.addmm(out, weight, x.reshape(fm_in, h * w), beta=0, alpha=1)      # :21
    out = out.reshape(1, weight.shape[0], h, w)
    return out + bias.reshape(1, -1, 1, 1).expand_as(out)                           # :26


def net_te2(x, layers):
    for i, (w, b) in enumerate(layers):                                           # main.lua:688-695
        x = conv1_fw(x, w, b)
        x = torch.relu(x) if i + 1 < len(layers) else torch.sigmoid(x)
    return x


def head_volume(output, D, layers, direction):
    """main.lua:962-978; output = (2, fm, H, W) tower outputs"""
    _, fm, H, W = output.shape
    vol = torch.full((1, D, H, W), float("nan"))
    for d in range(1, D + 1):                                                      # Lua's 1-based d
        l = output[0:1, :, :, d - 1:]                                              # :967  output[{{1},{},{},{d,-1}}]
        r = output[1:2, :, :, :W - d + 1]                                          # :968  output[{{2},{},{},{1,-d}}]
        x = torch.cat([l, r], dim=1)                                               # :969-972 (resize + two copies)
        o = net_te2(x, layers)[0, 0]                                               # :973
        if direction == -1:
            vol[0, d - 1, :, d - 1:] = o                                           # :976
        else:
            vol[0, d - 1, :, :W - d + 1] = o
    return vol


def main():
    rng = np.random.default_rng(7)
    fm, nh2, l2, H, W, D = 112, 384, 4, 3, 40, 12
    layers_np = sh.make_weights(rng, fm, nh2, l2)
    featL = np.maximum(rng.standard_normal((fm, H, W)), 0).astype(np.float32)     # the tower ends with a ReLU
    featR = np.maximum(rng.standard_normal((fm, H, W)), 0).astype(np.float32)
    layers = [(torch.from_numpy(w), torch.from_numpy(b)) for w, b in layers_np]
    out = torch.from_numpy(np.stack([featL, featR]))
    with torch.no_grad():
        volL = head_volume(out, D, layers, -1).numpy()
        volR = head_volume(out, D, layers, 1).numpy()
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scorer_head.npz")
    arrs = dict(featL=featL, featR=featR, volL=volL, volR=volR, D=np.int32(D))
    for i, (w, b) in enumerate(layers_np):
        arrs["w%d" % i] = w
        arrs["b%d" % i] = b
    np.savez_compressed(dst, **arrs)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
