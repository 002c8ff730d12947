"""CPU, property-based (hypothesis): the band decomposition of sgm2 that the multi-GPU row-band driver relies
on -- horizontal passes on ANY partition of the rows, then vertical passes on ANY partition of the columns,
accumulate exactly what the whole-image call accumulates -- and the host-side partition helper."""
import numpy as np
import pytest

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

import mccnn_b200  # noqa: E402,F401
from mccnn_b200 import rowband, synth  # noqa: E402


def _cuts(n, k, rng):
    """k-1 distinct interior cut points of [0, n) -> list of (lo, hi)"""
    k = min(k, n)
    pts = sorted(rng.choice(np.arange(1, n), size=k - 1, replace=False).tolist()) if k > 1 else []
    edges = [0] + pts + [n]
    return list(zip(edges[:-1], edges[1:]))


@settings(max_examples=25, deadline=None)
@given(H=st.integers(2, 9), W=st.integers(3, 24), D=st.integers(1, 40), direction=st.sampled_from([-1, 1]),
       nr=st.integers(1, 4), nc=st.integers(1, 4), seed=st.integers(0, 2 ** 16))
def test_sgm2_by_bands_equals_whole_image(oracle, H, W, D, direction, nr, nc, seed):
    hyp.assume(H <= W)                                                 # the reference's scratch layout needs it (main.lua:1012)
    rng = np.random.default_rng(seed)
    p = synth.make_pair(H, W, 3, D, seed=seed)
    volL, volR = oracle.stereo_join(p["featL"], p["featR"], D)
    vol = np.ascontiguousarray(oracle.transpose_dhw_to_hwd(volL if direction == -1 else volR))
    args = (1.3, 20.0, 0.08, 2.0, 3.0, 2.0, direction)
    want = oracle.sgm2(p["imgL"], p["imgR"], vol, *args)
    acc = np.zeros_like(vol)
    for y0, y1 in _cuts(H, nr, rng):                                   # passes 0, 1 (right, left) on row bands
        band = np.ascontiguousarray(acc[y0:y1])
        oracle.sgm2_band(p["imgL"], p["imgR"], np.ascontiguousarray(vol[y0:y1]), band, W, y0, 0, *args, 3)
        acc[y0:y1] = band
    for x0, x1 in _cuts(W, nc, rng):                                   # passes 2, 3 (down, up) on column bands
        band = np.ascontiguousarray(acc[:, x0:x1])
        oracle.sgm2_band(p["imgL"], p["imgR"], np.ascontiguousarray(vol[:, x0:x1]), band, W, 0, x0, *args, 12)
        acc[:, x0:x1] = band
    assert np.array_equal(acc, want, equal_nan=True)


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 5000), parts=st.integers(1, 64))
def test_split_is_a_partition(n, parts):
    pieces = [rowband.split(n, parts, i) for i in range(parts)]
    assert pieces[0][0] == 0 and pieces[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))      # contiguous, no gap, no overlap
    sizes = [hi - lo for lo, hi in pieces]
    assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1           # nearly equal
