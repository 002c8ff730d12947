"""GPU parity of StereoJoin's pitched fast path (8-byte loads, TMA store of the left volume) against the CPU oracle
(adcensus.cu:1455-1477).  Bar: bit-identical volumes; left volume NaN where x < d, right volume untouched where x >= W - d."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mccnn_b200  # noqa: E402,F401
from mccnn_b200 import adcensus, synth  # noqa: E402


def dev():
    return torch.device("cuda:0")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev())


@pytest.mark.parametrize("H,W,C,D", [
    (64, 128, 64, 16),    # BASELINE config 1; W % 4 == 0
    (7, 34, 5, 9),        # ragged, tiny, one tile
    (19, 150, 64, 70),
    (5, 300, 16, 228),    # two disparity chunks
    (3, 40, 128, 33),     # C at the reference limit
    (4, 262, 8, 120),     # D = one full chunk, W % 4 == 2
    (3, 130, 6, 150),     # D > W: rows that are entirely invalid
])
def test_stereo_join_pitched(oracle, H, W, C, D):
    p = synth.make_pair(H, W, C, D, seed=H + W)
    wantL, wantR = oracle.stereo_join(p["featL"], p["featR"], D)
    ld = (W + 3) // 4 * 4 + 4
    outL = torch.full((D, H, ld), 7.0, device=dev())
    outR = torch.full((D, H, ld), 9.0, device=dev())
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    fL, fR = cu(p["featL"]), cu(p["featR"])
    rc = adcensus.lib().mccnn_stereo_join_pitched(vp(fL), vp(fR), vp(outL), vp(outR), C, D, H, W, ld, adcensus._stream(fL))
    assert rc == 0, rc
    torch.cuda.synchronize()
    oL, oR = outL.cpu().numpy(), outR.cpu().numpy()
    # the padding columns hold unspecified values in the private layout: the left volume's TMA store clips at 16-byte
    # granularity, so the chunk that straddles W may receive the NaN of the x >= W mask; nothing else may appear there
    padL = oL[:, :, W:]
    assert (np.isnan(padL) | (padL == 7.0)).all() and (oR[:, :, W:] == 9.0).all(), "padding columns hold data"
    assert np.array_equal(oL[:, :, :W], wantL, equal_nan=True), "left volume (NaN where x < d)"
    wr = np.where(np.isnan(wantR), np.float32(9.0), wantR)       # untouched where x >= W - d
    assert np.array_equal(oR[:, :, :W], wr), "right volume"


def test_stereo_join_pitched_rejects_odd_width():
    t = torch.zeros((4, 3, 33), device=dev())
    o = torch.zeros((5, 3, 36), device=dev())
    vp = lambda x: ctypes.c_void_p(x.data_ptr())
    assert adcensus.lib().mccnn_stereo_join_pitched(vp(t), vp(t), vp(o), vp(o.clone()), 4, 5, 3, 33, 36, None) == -1
