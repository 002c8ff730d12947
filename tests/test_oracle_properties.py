"""CPU: the oracle against the reference's own (stale) unit-test definitions and the
invariants its compiled-out device asserts document."""
import numpy as np

import mccnn_b200  # noqa: F401
from mccnn_b200 import pipeline, synth


def test_stereojoin_matches_test_lua_definition(oracle):
    """test.lua:45-73: out[d,y,x] = sum_c L[c,y,x] * R[c,y,x-d], NaN where x-d < 0.
    The shipped kernel negates it (adcensus.cu:1470)."""
    rng = np.random.default_rng(0)
    L = rng.standard_normal((32, 10, 20)).astype(np.float32)
    R = rng.standard_normal((32, 10, 20)).astype(np.float32)
    outL, outR = oracle.stereo_join(L, R, 16)
    for d in range(16):
        for y in range(10):
            for x in range(20):
                if x - d < 0:
                    assert np.isnan(outL[d, y, x])
                else:
                    want = -float(np.dot(L[:, y, x].astype(np.float64), R[:, y, x - d].astype(np.float64)))
                    assert abs(outL[d, y, x] - want) < 1e-4
                    assert outR[d, y, x - d] == outL[d, y, x]
    for d in range(16):
        assert np.isnan(outR[d, :, 20 - d:]).all()


def test_normalize_matches_test_lua_definition(oracle):
    """test.lua:77-108 (the kernel adds 1e-5 under the root, adcensus.cu:1296)"""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 8, 5, 7)).astype(np.float32)
    out, norm = oracle.normalize_forward(x)
    want = x / np.sqrt((x.astype(np.float64) ** 2).sum(1, keepdims=True) + 1e-5)
    assert np.abs(out - want).max() < 1e-6
    assert np.abs((out.astype(np.float64) ** 2).sum(1) - 1).max() < 1e-4


def test_config1_pipeline_invariants(oracle):
    """BASELINE.json config 1 (64x128, d=16, CPU float tensors): whole chain on the oracle."""
    H, W, C, D = 64, 128, 64, 16
    p = synth.make_pair(H, W, C, D, seed=0)
    for preset in (("kitti", "fast"), ("kitti", "slow")):
        opt = pipeline.make_params(*preset)
        disp, vL, vR = oracle.stereo_predict(p["featL"], p["featR"], p["imgL"], p["imgR"], D,
                                             oracle.Params(**opt.as_dict()), want_vols=True)
        assert not np.isnan(disp).any()                      # main.lua:1224
        assert disp.min() >= 0 and disp.max() <= D - 1 + 1e-3
        for d in range(D):                                    # NaN triangles survive every stage
            assert np.isnan(vL[d, :, :d]).all() and not np.isnan(vL[d, :, d:]).any()
            assert np.isnan(vR[d, :, W - d:]).all() and not np.isnan(vR[d, :, :W - d]).any()
        assert (np.abs(disp - p["gt"]) < 1.0).mean() > 0.7   # recovers the synthetic ground truth


def test_cross_invariants(oracle):
    rng = np.random.default_rng(3)
    img = rng.standard_normal((20, 30)).astype(np.float32)
    for L1, tau in ((5, 0.5), (14, 0.2), (0, 0.0)):
        c = oracle.cross(img, L1, tau)
        xs = np.arange(30)[None, :]
        ys = np.arange(20)[:, None]
        assert (c[0] < xs).all() and (c[1] > xs).all() and (c[2] < ys).all() and (c[3] > ys).all()
        assert (c[0] >= -1).all() and (c[1] <= 30).all() and (c[2] >= -1).all() and (c[3] <= 20).all()
        lim = max(L1, 2)                                      # dist==1 always accepted (adcensus.cu:310)
        assert (xs - c[0] <= lim).all() and (c[1] - xs <= lim).all()


def test_cbca_of_constant_is_constant(oracle):
    H, W, D = 12, 25, 6
    img = np.random.default_rng(4).standard_normal((H, W)).astype(np.float32)
    x0c = oracle.cross(img, 5, 0.5)
    vol = np.full((D, H, W), np.nan, np.float32)
    for d in range(D):
        vol[d, :, d:] = 3.0
    out = oracle.cbca(x0c, x0c, vol, -1)
    assert np.array_equal(out, vol, equal_nan=True)           # cnt > 0, support never touches NaN (asserts :366,372)


def test_sgm_first_pixel_and_linearity_of_accumulation(oracle):
    H, W, D = 6, 14, 5
    rng = np.random.default_rng(5)
    vol = rng.random((H, W, D)).astype(np.float32)
    img = rng.standard_normal((H, W)).astype(np.float32)
    args = (1.0, 8.0, 0.1, 2.0, 3.0, 2.0, -1)
    out = oracle.sgm2(img, img, vol, *args)
    init = rng.random((H, W, D)).astype(np.float32)
    out2 = oracle.sgm2(img, img, vol, *args, out=init.copy())
    assert np.abs((out2 - init) - out).max() < 1e-5           # out += val (adcensus.cu:569,616)
    assert (out >= 4 * vol.min() - 1e-5).all()
    # with zero penalties every path cost equals the input cost: out = 4 * vol exactly
    out0 = oracle.sgm2(img, img, vol, 0.0, 0.0, 0.1, 2.0, 3.0, 2.0, -1)
    mn = vol.min(axis=2, keepdims=True)
    assert out0.shape == vol.shape and np.isfinite(out0).all() and (out0 <= 4 * vol + 1e-5).all() and (out0 >= 4 * mn - 1e-5).all()


def test_median_and_mean_border_handling(oracle):
    img = np.arange(20, dtype=np.float32).reshape(4, 5)
    m = oracle.median2d(img, 5)
    # corner (0,0): taps rows 0..2 x cols 0..2 = 9 values -> xs[4] (adcensus.cu:1592)
    assert m[0, 0] == np.sort(img[:3, :3].ravel())[4]
    k = oracle.gaussian(0.5)
    assert k.shape == (5, 5) and k[2, 2] == 1.0
    out = oracle.mean2d(img, k, 1000.0)
    assert np.isfinite(out).all() and abs(out[2, 2] - img[2, 2]) < 1.0


def test_chain_composition_equals_monolith(oracle):
    """oracle.stereo_predict_chain (per-operator composition, also the ad / census vehicle) must give
    exactly what the one-call C chain gives for arch 'fast'."""
    H, W, C, D = 24, 60, 8, 12
    p = synth.make_pair(H, W, C, D, seed=5)
    for preset in (("kitti", "fast"), ("kitti", "accurate_cbca4"), ("mb", "fast")):
        prm = oracle.Params(**pipeline.make_params(*preset).as_dict())
        a = oracle.stereo_predict(p["featL"], p["featR"], p["imgL"], p["imgR"], D, prm, want_vols=True)
        b = oracle.stereo_predict_chain(p["imgL"], p["imgR"], D, prm, "fast", p["featL"], p["featR"], want_vols=True)
        for x, y in zip(a, b):
            assert np.array_equal(x, y, equal_nan=True), preset


def test_ad_census_chain_invariants(oracle):
    """main.lua:932-942: net-free volumes.  NaN triangles as for StereoJoin, census costs are counts
    of differing bits / channel in [0, 81], the chain returns a finite map in range and recovers a
    constant shift of a textured image."""
    H, W, D, shift = 40, 96, 12, 5
    rng = np.random.default_rng(7)
    base = synth.natural_image(rng, H, W + shift)
    base = ((base - base.mean()) / base.std(ddof=1)).astype(np.float32)   # one gain/offset for both views (AD is not invariant)
    imgL, imgR = base[:, :W].copy(), base[:, shift:].copy()      # left pixel x shows what right pixel x - shift shows
    vc = oracle.census(imgL, imgR, D, -1)
    va = oracle.ad(imgL, imgR, D, -1)
    for d in range(D):
        for v in (vc, va):
            assert np.isnan(v[d, :, :d]).all() and not np.isnan(v[d, :, d:]).any()
    assert vc[~np.isnan(vc)].min() >= 0 and vc[~np.isnan(vc)].max() <= 81 and np.all(vc[~np.isnan(vc)] == np.round(vc[~np.isnan(vc)]))
    assert va[~np.isnan(va)].min() >= 0
    for arch in ("census", "ad"):
        prm = oracle.Params(**pipeline.make_params("kitti", arch).as_dict())
        disp = oracle.stereo_predict_chain(imgL, imgR, D, prm, arch)
        assert disp.shape == (H, W) and not np.isnan(disp).any()
        assert disp.min() >= 0 and disp.max() <= D - 1 + 1e-3
        assert (np.abs(disp[:, 2 * D:] - shift) < 1.0).mean() > 0.9, arch


def test_chain_from_given_volumes_equals_the_fast_arch_chain(oracle):
    """arch 'volumes' (the accurate architecture's entry: the scorer head's volumes are given, main.lua:981 onwards) is the same
    chain as arch 'fast' after StereoJoin"""
    import mccnn_b200  # noqa: F401
    from mccnn_b200 import pipeline, synth

    H, W, C, D = 20, 44, 4, 9
    p = synth.make_pair(H, W, C, D, seed=21)
    opt = oracle.Params(**pipeline.make_params("kitti", "slow").as_dict())
    a = oracle.stereo_predict_chain(p["imgL"], p["imgR"], D, opt, arch="fast", featL=p["featL"], featR=p["featR"])
    vl, vr = oracle.stereo_join(p["featL"], p["featR"], D)
    b = oracle.stereo_predict_chain(p["imgL"], p["imgR"], D, opt, arch="volumes", volL=vl, volR=vr)
    assert np.array_equal(a, b, equal_nan=True)
