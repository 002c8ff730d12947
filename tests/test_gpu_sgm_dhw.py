"""GPU parity of sgm2 on the (D, H, ld) layout (csrc/sgm_dhw.cu) against the CPU oracle's sgm2 on the reference's
(H, W, D) layout (adcensus.cu:535-697).  Bar: bit-identical (adds / fminf / IEEE divisions only, directions
accumulated in the reference's order), including the /4 of main.lua:1020."""
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


@pytest.mark.parametrize("H,W,D,direction,div4", [
    (11, 37, 20, -1, 1),     # K = 1, ragged width (W % 4 = 1)
    (10, 45, 70, 1, 0),      # K = 4
    (9, 50, 40, -1, 1),      # K = 2
    (7, 300, 228, -1, 1),    # K = 8, the bench's disparity range
    (6, 310, 300, 1, 1),     # K = 16
    (33, 64, 16, -1, 1),     # W % 4 == 0
    (5, 6, 3, 1, 0),         # a single partial column group pair
    (4, 5, 2, -1, 1),        # two disparities, one column group pair (H > W cannot be checked: the oracle keeps the reference's
                             # (W, D) line-state scratch, whose horizontal passes alias for H > W -- and race under OpenMP)
])
def test_sgm2_dhw(oracle, H, W, D, direction, div4):
    p = synth.make_pair(H, W, 4, D, seed=D + H)
    volL, volR = oracle.stereo_join(p["featL"], p["featR"], D)
    vol = volL if direction == -1 else volR
    args = (1.32, 24.25, 0.08, 2.0, 3.0, 2.0)
    assert H <= W
    want = oracle.sgm2(p["imgL"], p["imgR"], oracle.transpose_dhw_to_hwd(vol), *args, direction)
    want = np.ascontiguousarray(want.transpose(2, 0, 1))
    if div4:
        want = want / np.float32(4)
    ld = (W + 3) // 4 * 4 + 4
    vin = torch.full((D, H, ld), 777.0, device=dev())
    vin[:, :, :W] = cu(vol)
    acc = torch.full((D, H, ld), float("nan"), device=dev())     # need not be initialised
    vp = lambda t_: ctypes.c_void_p(t_.data_ptr())
    cf = ctypes.c_float
    iL, iR = cu(p["imgL"]), cu(p["imgR"])
    rc = adcensus.lib().mccnn_sgm2_dhw(vp(iL), vp(iR), vp(vin), vp(acc), H, W, ld, D, cf(args[0]), cf(args[1]), cf(args[2]),
                                       cf(args[3]), cf(args[4]), cf(args[5]), direction, div4, adcensus._stream(vin))
    assert rc == 0, rc
    torch.cuda.synchronize()
    got = acc[:, :, :W].cpu().numpy()
    if not np.array_equal(got, want, equal_nan=True):
        bad = ~((got == want) | (np.isnan(got) & np.isnan(want)))
        raise AssertionError("%d / %d differ, first %s got %s want %s" % (bad.sum(), bad.size, np.argwhere(bad)[:4].tolist(),
                                                                          got[bad][:4], want[bad][:4]))
