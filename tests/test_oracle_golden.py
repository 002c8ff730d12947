"""CPU: pin the oracle (oracle/adcensus_oracle.c) against the golden fixtures, i.e.
against outputs of the reference's own kernels (tests/golden/README.md).  Bit-exact,
stage by stage along main.lua's stereo_predict chain."""
import glob
import os

import numpy as np
import pytest


def same(a, b, what):
    a = np.asarray(a).reshape(np.asarray(b).shape)
    if not np.array_equal(a, b, equal_nan=True):
        bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
        raise AssertionError("%s: %d / %d differ; got %s want %s" % (what, bad.sum(), bad.size, a[bad][:4], b[bad][:4]))


def load_opt(g, oracle):
    kw = {}
    for k, v in zip(g["opt_names"], g["opt_values"]):
        kw[str(k)] = float(v) if str(k) in ("tau1", "pi1", "pi2", "sgm_q1", "sgm_q2", "alpha1", "tau_so", "blur_sigma",
                                            "blur_t") else int(v)
    return oracle.Params(**kw)


def pipe_files(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "pipe_*.npz")))


def test_fixtures_present(golden_dir):
    assert len(pipe_files(golden_dir)) >= 3 and os.path.exists(os.path.join(golden_dir, "ops.npz"))


@pytest.mark.parametrize("name", ["pipe_kitti_slow", "pipe_kitti_fast", "pipe_mb_slow"])
def test_stage_by_stage(oracle, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    H, W, C, D = [int(v) for v in g["meta"]]
    opt = load_opt(g, oracle)
    fL, fR, iL, iR = g["featL"], g["featR"], g["imgL"], g["imgR"]

    sjL, sjR = oracle.stereo_join(fL, fR, D)
    same(sjL, g["sj_left"], "StereoJoin left")
    same(sjR, g["sj_right"], "StereoJoin right")
    x0c = oracle.cross(iL, opt.L1, opt.tau1)
    x1c = oracle.cross(iR, opt.L1, opt.tau1)
    same(x0c, g["x0c"], "cross left")
    same(x1c, g["x1c"], "cross right")
    vols = {-1: oracle.fix_border(sjL.copy(), opt.border, -1), 1: oracle.fix_border(sjR.copy(), opt.border, 1)}
    disp = {}
    for direction, tag in ((1, "R"), (-1, "L")):
        vol = vols[direction]
        for _ in range(opt.cbca_i1):
            vol = oracle.cbca(x0c, x1c, vol, direction)
        same(vol, g["cbca1_" + tag], "cbca1 " + tag)
        out = oracle.sgm2(iL, iR, oracle.transpose_dhw_to_hwd(vol), opt.pi1, opt.pi2, opt.tau_so, opt.alpha1,
                          opt.sgm_q1, opt.sgm_q2, direction)
        vol = oracle.transpose_hwd_to_dhw_div4(out)
        same(vol, g["sgm_" + tag], "sgm " + tag)
        for _ in range(opt.cbca_i2):
            vol = oracle.cbca(x0c, x1c, vol, direction)
        vols[direction] = vol
        disp[direction] = oracle.spatial_argmin(vol) - 1
    same(disp[1], g["disp_R"], "argmin right")
    same(disp[-1], g["disp_L"], "argmin left")
    d = disp[-1]
    if opt.lr_check:
        outlier = oracle.outlier_detection(disp[-1], disp[1], D)
        same(outlier, g["outlier"], "outlier_detection")
        d = oracle.interpolate_occlusion(d, outlier)
        same(d, g["occ"], "interpolate_occlusion")
        d = oracle.interpolate_mismatch(d, outlier)
        same(d, g["mis"], "interpolate_mismatch")
    d = oracle.subpixel_enchancement(d, vols[-1], D)
    same(d, g["subpixel"], "subpixel_enchancement")
    d = oracle.median2d(d, 5)
    same(d, g["median"], "median2d")
    d = oracle.mean2d(d, oracle.gaussian(opt.blur_sigma), opt.blur_t)
    same(d, g["disp"], "mean2d / disp")

    # and the orchestration restatement end to end
    full = oracle.stereo_predict(fL, fR, iL, iR, D, opt)
    same(full, g["disp"], "orc_stereo_predict")


def test_standalone_ops(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "ops.npz"))
    out, norm = oracle.normalize_forward(g["norm_in"])
    same(out, g["norm_out"], "Normalize_forward")
    same(norm, g["norm_norm"], "Normalize_forward norm")
    a, b = g["adc_img0"][0, 0], g["adc_img1"][0, 0]
    for direction in (-1, 1):
        same(oracle.ad(a, b, 7, direction), g["ad_%d" % direction], "ad")
        same(oracle.census(a[None], b[None], 7, direction), g["census_%d" % direction], "census")
    vin = g["argmin_in"]
    for n in range(vin.shape[0]):
        same(oracle.spatial_argmin(vin[n]), g["argmin_out"][n, 0], "spatial_argmin")
    img = g["post_img"][0, 0]
    for k in (3, 5, 7):
        same(oracle.median2d(img, k), g["median_%d" % k], "median2d k=%d" % k)
    same(oracle.mean2d(img, g["mean2d_kernel"], 3.0), g["mean2d_out"], "mean2d")
    same(oracle.gaussian(1.2), g["mean2d_kernel"], "gaussian")
