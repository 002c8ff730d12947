"""GPU parity of the fused tcgen05 scorer head (csrc/scorer_head.cu) against the CPU oracle (oracle/scorer_head.py, pinned
to the PyTorch restatement of main.lua:958-984 / SpatialConvolution1_fw.lua by tests/golden/scorer_head.npz).
Bar: the north star's 1e-4 for float work (output is a sigmoid in (0, 1): absolute), NaN pattern identical."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mccnn_b200  # noqa: E402,F401
from mccnn_b200 import scorer_head  # noqa: E402
from oracle import scorer_head as osh  # noqa: E402


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to("cuda:0")


def check(got, want, bar):
    got = got.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern differs"
    m = ~np.isnan(want)
    err = float(np.abs(got[m] - want[m]).max())
    assert err <= bar, "max abs error %.3g (bar %.1g)" % (err, bar)
    return err


def test_golden_fixture():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "scorer_head.npz"))
    layers = [(g["w%d" % i], g["b%d" % i]) for i in range(5)]
    head = scorer_head.ScorerHead(layers)
    vL, vR = head.volumes(cu(g["featL"]), cu(g["featR"]), int(g["D"]))
    torch.cuda.synchronize()
    check(vL, g["volL"], 1e-4)
    check(vR, g["volR"], 1e-4)
    head.close()


@pytest.mark.parametrize("fm,nh2,l2,H,W,D", [
    (112, 384, 4, 5, 300, 40),     # kitti net, three tiles per row, ragged last tile
    (112, 384, 3, 2, 130, 130),    # kitti2015 / mb depth (l2 = 3), D reaches past the first tile
    (64, 256, 2, 3, 128, 9),       # one 256-wide instruction per step
    (8, 128, 1, 2, 77, 5),         # smallest legal shapes
])
def test_against_oracle(fm, nh2, l2, H, W, D):
    rng = np.random.default_rng(fm + l2)
    layers = osh.make_weights(rng, fm, nh2, l2)
    fL = np.maximum(rng.standard_normal((fm, H, W)), 0).astype(np.float32)
    fR = np.maximum(rng.standard_normal((fm, H, W)), 0).astype(np.float32)
    head = scorer_head.ScorerHead(layers)
    vL, vR = head.volumes(cu(fL), cu(fR), D)
    torch.cuda.synchronize()
    wantL = osh.head_volume(fL, fR, D, layers, -1)[None]
    wantR = osh.head_volume(fL, fR, D, layers, 1)[None]
    e = check(vL, wantL, 1e-4)
    check(vR, wantR, 1e-4)
    assert e < 2e-5, "the bf16-split path should sit two orders inside the bar, got %.3g" % e
    # plain bf16 operands (nterms = 1): the fast, not fp32-grade mode
    v1, _ = head.volumes(cu(fL), cu(fR), D, nterms=1, want_right=False)
    torch.cuda.synchronize()
    check(v1, wantL, 2e-2)
    head.close()


def test_rejects_unsupported_shapes():
    rng = np.random.default_rng(0)
    with pytest.raises(Exception):
        scorer_head.ScorerHead(osh.make_weights(rng, 8, 100, 2))       # nh2 not a multiple of 128
