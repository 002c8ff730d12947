"""GPU parity of the constant-work, TMA-staged CBCA (csrc/cbca_tma.cu) on pitched volumes against the CPU
oracle's tap-by-tap sums (adcensus.cu:343-377).  Bar: the north star's 1e-4 relative for float aggregation,
NaN positions identical."""
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


def pitched(vol, ld, fill=0.0):
    """(D,H,W) numpy -> (D,H,ld) device tensor, padding columns hold `fill`"""
    D, H, W = vol.shape
    t = torch.full((D, H, ld), fill, device=dev(), dtype=torch.float32)
    t[:, :, :W] = cu(vol)
    return t


def run_fast(x0c, x1c, vol, direction, max_arm, ld):
    D, H, W = vol.shape
    vin = pitched(vol, ld, fill=123.0)      # the padding must never be read as data
    vout = torch.full((D, H, ld), -7.0, device=dev())
    a0, a1 = cu(x0c), cu(x1c)
    vp = lambda t_: ctypes.c_void_p(t_.data_ptr())
    rc = adcensus.lib().mccnn_cbca_fast_pitched(vp(a0), vp(a1), vp(vin), vp(vout), D, H, W, ld, direction, max_arm,
                                                adcensus._stream(vin))
    assert rc == 0, rc
    torch.cuda.synchronize()
    o = vout.cpu().numpy()
    assert (o[:, :, W:] == -7.0).all(), "padding columns were written"
    return o[:, :, :W]


def check(got, want):
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern differs"
    m = ~np.isnan(want)
    err = np.abs(got[m] - want[m]) / np.maximum(1.0, np.abs(want[m]))
    assert err.max() <= 1e-4, "max relative error %.3g" % err.max()
    return float(err.max())


@pytest.mark.parametrize("H,W,D,L1,tau1,direction", [
    (70, 300, 36, 5, 0.13, -1),
    (70, 300, 36, 5, 0.13, 1),
    (33, 130, 20, 2, 0.5, 1),       # halo 1
    (64, 128, 16, 5, 5.0, -1),      # every arm at full length
    (45, 270, 30, 9, 5.0, -1),      # halo 8
    (50, 300, 40, 14, 0.02, -1),    # Middlebury preset arms, halo 13
    (40, 140, 30, 14, 5.0, 1),      # halo 13, all arms full
    (7, 40, 9, 5, 0.13, -1),        # smaller than one tile
    (37, 257, 150, 5, 0.13, -1),    # D chunks entirely inside the invalid triangle
])
def test_cbca_tma_vs_oracle(oracle, H, W, D, L1, tau1, direction):
    C = 8
    p = synth.make_pair(H, W, C, D, seed=L1 + H)
    volL, volR = oracle.stereo_join(p["featL"], p["featR"], D)
    vol = volL if direction == -1 else volR
    x0c, x1c = oracle.cross(p["imgL"], L1, tau1), oracle.cross(p["imgR"], L1, tau1)
    want = oracle.cbca(x0c, x1c, vol, direction)
    ld = (W + 3) // 4 * 4 + 4
    got = run_fast(x0c, x1c, vol, direction, max(L1, 2), ld)
    check(got, want)


def test_cbca_tma_four_iterations(oracle):
    """error growth over the CBCA x 4 of the bench preset stays far inside the bar"""
    H, W, D, L1, tau1 = 64, 200, 24, 5, 0.13
    p = synth.make_pair(H, W, 8, D, seed=11)
    volL, _ = oracle.stereo_join(p["featL"], p["featR"], D)
    x0c, x1c = oracle.cross(p["imgL"], L1, tau1), oracle.cross(p["imgR"], L1, tau1)
    want, got = volL, volL
    for _ in range(4):
        want = oracle.cbca(x0c, x1c, want, -1)
        got = run_fast(x0c, x1c, got, -1, L1, 200)
    assert check(got, want) < 2e-5


def test_cbca_tma_rejects_bad_pitch(oracle):
    t = torch.zeros((4, 8, 30), device=dev())
    a = torch.zeros((4, 8, 30), device=dev())
    vp = lambda t_: ctypes.c_void_p(t_.data_ptr())
    rc = adcensus.lib().mccnn_cbca_fast_pitched(vp(a), vp(a), vp(t), vp(t.clone()), 4, 8, 30, 30, -1, 5, None)
    assert rc == -1
