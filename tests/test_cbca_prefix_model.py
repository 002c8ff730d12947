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


def _pack_arms(xc):
    """pack_arms_kernel (cross_cbca.cu): arm LENGTHS as bytes, L | R << 8 | U << 16 | D << 24"""
    H, W = xc.shape[1:]
    ys, xs = np.mgrid[0:H, 0:W]
    l = np.clip(xs - xc[0], 0, 255).astype(np.uint32)
    r = np.clip(xc[1] - xs, 0, 255).astype(np.uint32)
    u = np.clip(ys - xc[2], 0, 255).astype(np.uint32)
    d = np.clip(xc[3] - ys, 0, 255).astype(np.uint32)
    return l | (r << 8) | (u << 16) | (d << 24)


def _emulate_o1_kernel(a0, a1, vol, direction, R, TX=128, TY=32):
    """cbca_o1_kernel (cross_cbca.cu) restated index for index in numpy: same tile, halo, offsets and buffers"""
    D, H, W = vol.shape
    TH, TWP = TY + 2 * R + 1, TX + 2 * R + 1               # one extra row / column: the excluded end-points are indexed
    out = np.full_like(vol, np.float32(-12345.0))
    bmin = lambda p, q: np.minimum(p & 255, q & 255) | (np.minimum((p >> 8) & 255, (q >> 8) & 255) << 8) | \
        (np.minimum((p >> 16) & 255, (q >> 16) & 255) << 16) | (np.minimum(p >> 24, q >> 24) << 24)
    for d in range(D):
        sh = d * direction
        for y0 in range(0, H, TY):
            for x0 in range(0, W, TX):
                scomb = np.zeros((TH, TX), np.uint32)
                sv = np.zeros((TH, TWP), np.float32)
                for r in range(TH):
                    yy = y0 - R - 1 + r
                    if not (0 <= yy < H):
                        continue
                    for c in range(TX):
                        xx, xr = x0 + c, x0 + c + sh
                        a = a0[yy, xx] if xx < W else np.uint32(0)
                        b = a1[yy, xr] if 0 <= xr < W else np.uint32(0)
                        scomb[r, c] = bmin(np.uint32(a), np.uint32(b))
                    for c in range(TWP):
                        xx = x0 - R - 1 + c
                        if 0 <= xx < W:
                            v = vol[d, yy, xx]
                            sv[r, c] = v if v == v else 0.0
                I = np.cumsum(sv, axis=1, dtype=np.float32)                       # 1. row prefixes, in place
                sS = np.zeros((TH, TX), np.float32)
                sN = np.zeros((TH, TX), np.float32)
                for r in range(TH):                                              # 2. run sums / lengths
                    for cx in range(TX):
                        c = int(scomb[r, cx])
                        L, Rr = c & 255, (c >> 8) & 255
                        if L > 0:
                            sS[r, cx] = I[r, cx + R + 1 + Rr - 1] - I[r, cx + R + 1 - L]
                            sN[r, cx] = L + Rr - 1
                sS = np.cumsum(sS, axis=0, dtype=np.float32)                      # 3. column prefixes
                sN = np.cumsum(sN, axis=0, dtype=np.float32)
                for ty_ in range(TY):                                            # 4. outputs
                    y = y0 + ty_
                    if y >= H:
                        continue
                    for cx in range(TX):
                        x = x0 + cx
                        if x >= W:
                            continue
                        xs = x + sh
                        if 0 <= xs < W:
                            ty = ty_ + R + 1
                            c = int(scomb[ty, cx])
                            U, Dn = (c >> 16) & 255, c >> 24
                            hi, lo = ty + Dn - 1, ty - U
                            with np.errstate(invalid="ignore", divide="ignore"):
                                out[d, y, x] = (sS[hi, cx] - sS[lo, cx]) / (sN[hi, cx] - sN[lo, cx])
                        else:
                            out[d, y, x] = vol[d, y, x]
    return out


@pytest.mark.parametrize("L1,tau1,direction,H,W,D", [(5, 0.13, -1, 37, 140, 7), (5, 0.13, 1, 37, 140, 7),
                                                      (2, 0.5, -1, 20, 131, 5), (9, 5.0, 1, 34, 129, 4)])
def test_o1_kernel_indexing_emulated(oracle, L1, tau1, direction, H, W, D):
    """the experimental kernel's tile / halo / offset arithmetic, emulated index for index on the CPU (one and a
    bit tiles in x and y, image edges, both directions), against the reference's tap sums"""
    x0c, x1c, volL, volR = _case(oracle, H, W, D, L1, tau1, seed=H + L1)
    vol = volL if direction == -1 else volR
    want = oracle.cbca(x0c, x1c, vol, direction)
    maxlen = max(L1, 2)
    R = 1 if maxlen - 1 <= 1 else 4 if maxlen - 1 <= 4 else 8 if maxlen - 1 <= 8 else 13   # adc_cbca_packed's choice
    got = _emulate_o1_kernel(_pack_arms(x0c), _pack_arms(x1c), vol, direction, R)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    m = ~np.isnan(want)
    assert (np.abs(got[m] - want[m]) / np.maximum(1.0, np.abs(want[m]))).max() < 1e-4 / 4
