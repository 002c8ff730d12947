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


def test_arch_slow_chain_from_images(oracle):
    """main.lua's arch 'slow' flow end to end on the GPU: images -> feature tower (tcgen05) -> scorer head (tcgen05) ->
    fix_border -> cross / cbca / sgm2 / post (pipeline.stereo_predict(arch='slow')).  The stages after the head are integer /
    exact-order work: from the head's own volumes they must equal the oracle chain bit for bit."""
    from mccnn_b200 import feature_tower, pipeline, synth
    from oracle import feature_tower as oft

    H, W, D, fm, nh2 = 40, 150, 24, 16, 128
    rng = np.random.default_rng(3)
    p = synth.make_pair(H, W, 4, D, seed=9)
    x = np.stack([p["imgL"], p["imgR"]])[:, None].astype(np.float32)
    tower = feature_tower.FeatureTower(oft.make_weights(rng, l1=3, fm=fm), arch="slow")
    head = scorer_head.ScorerHead(osh.make_weights(rng, fm, nh2, 2))
    opt = pipeline.make_params("kitti", "slow")
    xb = cu(x)
    feats = tower.forward(xb)
    disp, vl, vr = pipeline.stereo_predict(xb, feats, opt, D, want_vols=True, arch="slow", head=head)
    hl, hr = head.volumes(feats[0].contiguous(), feats[1].contiguous(), D)
    torch.cuda.synchronize()
    want, wl, wr = oracle.stereo_predict_chain(p["imgL"], p["imgR"], D, oracle.Params(**opt.as_dict()), arch="volumes",
                                               volL=hl[0].cpu().numpy(), volR=hr[0].cpu().numpy(), want_vols=True)
    for got, ref, what in ((vl, wl, "left volume"), (vr, wr, "right volume"), (disp, want, "disparity map")):
        g = got.cpu().numpy().reshape(ref.shape)
        assert np.array_equal(g, ref, equal_nan=True), what + " differs from the oracle chain"
    tower.close()
    head.close()
