"""GPU parity of the in-library feature tower (csrc/feature_tower.cu) against the CPU oracle (oracle/feature_tower.py, pinned to
the PyTorch restatement of main.lua:726-749 by tests/golden/feature_tower.npz).  Bar: 1e-4 (unit-norm features: absolute)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mccnn_b200  # noqa: E402,F401
from mccnn_b200 import feature_tower  # noqa: E402
from oracle import feature_tower as oft  # noqa: E402


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to("cuda:0")


def test_golden_fixture():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "feature_tower.npz"))
    layers = [(g["w%d" % i], g["b%d" % i]) for i in range(4)]
    t = feature_tower.FeatureTower(layers, arch="fast")
    got = t.forward(cu(g["x"])).cpu().numpy()
    assert np.abs(got - g["out"]).max() <= 1e-4
    assert np.abs(got - g["out"]).max() <= 2e-5, "the bf16-split path should sit well inside the bar"
    t.close()


@pytest.mark.parametrize("n_in,fm,l1,arch,H,W", [
    (1, 64, 4, "fast", 9, 300),      # kitti fast: three tiles per row, ragged last tile
    (1, 64, 5, "fast", 5, 129),      # mb fast depth
    (1, 112, 4, "slow", 4, 200),     # accurate arch tower: 112 planes (padded to 128), ReLU after the last layer, no Normalize
    (3, 16, 2, "slow", 6, 70),       # colour input, small planes
])
def test_against_oracle(n_in, fm, l1, arch, H, W):
    rng = np.random.default_rng(fm + l1)
    layers = oft.make_weights(rng, l1=l1, fm=fm, n_in=n_in)
    x = rng.standard_normal((2, n_in, H, W)).astype(np.float32)
    want = oft.tower_forward(x, layers, normalize=(arch == "fast"))
    if arch == "slow":
        want = np.maximum(want, 0)                                   # main.lua:684-685: ReLU after every layer
    t = feature_tower.FeatureTower(layers, arch=arch)
    got = t.forward(cu(x)).cpu().numpy()
    scale = max(1.0, float(np.abs(want).max()))
    assert np.abs(got - want).max() <= 1e-4 * scale
    got1 = t.forward(cu(x), nterms=1).cpu().numpy()
    assert np.abs(got1 - want).max() <= 5e-2 * scale                 # plain bf16 operands
    t.close()
