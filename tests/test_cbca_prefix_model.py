"""CPU: design study for a tolerance-mode CBCA -- the prefix-sum formulation (constant work per pixel) must
reproduce the reference's tap-by-tap sums within rounding, with the error bounded by the tile size."""
import numpy as np
import pytest

import mccnn_b200  # noqa: F401
from mccnn_b200 import synth
from oracle import cbca_prefix_model as model


def _case(oracle, H, W, D, L1, tau1, seed):
    p = synth.make_pair(H, W, 8, D, seed=seed)
    volL, volR = oracle.stereo_join(p["featL"], p["featR"], D)
    return oracle.cross(p["imgL"], L1, tau1), oracle.cross(p["imgR"], L1, tau1), volL, volR


@pytest.mark.parametrize("L1,tau1", [(5, 0.13), (14, 0.02), (0, 0.0)])
@pytest.mark.parametrize("direction", [-1, 1])
def test_prefix_formulation_matches_the_tap_sums(oracle, L1, tau1, direction):
    x0c, x1c, volL, volR = _case(oracle, 40, 150, 12, L1, tau1, seed=L1 + 3)
    vol = volL if direction == -1 else volR
    want = oracle.cbca(x0c, x1c, vol, direction)
    got64 = model.cbca_prefix(x0c, x1c, vol, direction, dtype=np.float64)
    assert np.array_equal(np.isnan(got64), np.isnan(want))
    m = ~np.isnan(want)
    rel = np.abs(got64[m] - want[m]) / np.maximum(1.0, np.abs(want[m]))
    assert rel.max() < 2e-6                                   # same sums up to fp32 rounding of the reference itself
    got32 = model.cbca_prefix(x0c, x1c, vol, direction, dtype=np.float32, tile=(32, 128))
    rel32 = np.abs(got32[m] - want[m]) / np.maximum(1.0, np.abs(want[m]))
    assert rel32.max() < 1e-4 / 4                             # fp32 prefixes on 32x128 tiles: well inside the 1e-4 bar
    # the decision it feeds: arg-min over d (first minimum, NaN skipped) almost never changes
    am_w = np.nanargmin(np.where(np.isnan(want), np.inf, want), axis=0)
    am_g = np.nanargmin(np.where(np.isnan(got32), np.inf, got32), axis=0)
    assert (am_w != am_g).mean() < 5e-3


def test_error_grows_with_prefix_length(oracle):
    """why the kernel must keep prefixes tile-local: whole-row fp32 prefixes lose more digits"""
    x0c, x1c, volL, _ = _case(oracle, 24, 600, 6, 5, 0.13, seed=11)
    want = oracle.cbca(x0c, x1c, volL, -1)
    m = ~np.isnan(want)
    err = {}
    for tile in ((24, 64), (24, 600)):
        got = model.cbca_prefix(x0c, x1c, volL, -1, dtype=np.float32, tile=tile)
        err[tile] = (np.abs(got[m] - want[m]) / np.maximum(1.0, np.abs(want[m]))).max()
    assert err[(24, 64)] <= err[(24, 600)] * 1.5 + 1e-9 and err[(24, 600)] < 1e-4
