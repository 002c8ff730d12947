"""ctypes/numpy face of oracle/liboracle.so -- TEST INFRASTRUCTURE, not product code.

The CPU restatement of the reference's stereo-method operators lives in
adcensus_oracle.c (each function cites the adcensus.cu / main.lua lines it
follows).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import this module; the product (mc-cnn_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

f32p = ctypes.POINTER(ctypes.c_float)


class Params(ctypes.Structure):
    """Mirror of orc_params (adcensus_oracle.h); field names follow main.lua's options."""

    _fields_ = [
        ("L1", ctypes.c_int),
        ("tau1", ctypes.c_float),
        ("cbca_i1", ctypes.c_int),
        ("cbca_i2", ctypes.c_int),
        ("pi1", ctypes.c_float),
        ("pi2", ctypes.c_float),
        ("sgm_q1", ctypes.c_float),
        ("sgm_q2", ctypes.c_float),
        ("alpha1", ctypes.c_float),
        ("tau_so", ctypes.c_float),
        ("sgm_i", ctypes.c_int),
        ("blur_sigma", ctypes.c_double),
        ("blur_t", ctypes.c_float),
        ("border", ctypes.c_int),
        ("lr_check", ctypes.c_int),
    ]


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "adcensus_oracle.c"))
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _p(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(f32p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def normalize_forward(x):
    x = _f(x)
    N, C, H, W = x.shape
    norm = np.empty((N, 1, H, W), np.float32)
    out = np.empty_like(x)
    lib().orc_normalize_forward(_p(x), _p(norm), _p(out), N, C, H, W)
    return out, norm


def stereo_join(L, R, D, outL=None, outR=None):
    """L, R: (C,H,W).  Returns (outL, outR) of shape (D,H,W), NaN pre-filled as main.lua:946."""
    L, R = _f(L), _f(R)
    C, H, W = L.shape
    if outL is None:
        outL = np.full((D, H, W), np.nan, np.float32)
    if outR is None:
        outR = np.full((D, H, W), np.nan, np.float32)
    lib().orc_stereo_join(_p(L), _p(R), _p(outL), _p(outR), C, D, H, W)
    return outL, outR


def fix_border(vol, n, direction):
    D, H, W = vol.shape
    lib().orc_fix_border(_p(vol), D, H, W, n, direction)
    return vol


def ad(x0, x1, D, direction):
    x0, x1 = _f(x0), _f(x1)
    H, W = x0.shape[-2:]
    out = np.empty((D, H, W), np.float32)
    lib().orc_ad(_p(x0), _p(x1), _p(out), D, H, W, direction)
    return out


def census(x0, x1, D, direction):
    x0, x1 = _f(x0), _f(x1)
    nch = x0.shape[0] if x0.ndim == 3 else 1
    H, W = x0.shape[-2:]
    out = np.empty((D, H, W), np.float32)
    lib().orc_census(_p(x0), _p(x1), _p(out), D, nch, H, W, direction)
    return out


def cross(img, L1, tau1):
    img = _f(img)
    H, W = img.shape[-2:]
    out = np.empty((4, H, W), np.float32)
    lib().orc_cross(_p(img), _p(out), H, W, int(L1), ctypes.c_float(tau1))
    return out


def cbca(x0c, x1c, vol, direction):
    x0c, x1c, vol = _f(x0c), _f(x1c), _f(vol)
    D, H, W = vol.shape
    out = np.empty_like(vol)
    lib().orc_cbca(_p(x0c), _p(x1c), _p(vol), _p(out), D, H, W, direction)
    return out


def sgm2(x0, x1, vol_hwd, pi1, pi2, tau_so, alpha1, sgm_q1, sgm_q2, direction, out=None):
    """vol_hwd: (H,W,D).  Accumulates into `out` (zeros if None) and returns it."""
    x0, x1, vol_hwd = _f(x0), _f(x1), _f(vol_hwd)
    H, W, D = vol_hwd.shape
    if H > W:
        raise ValueError("sgm2: the reference's line-state scratch is (W, D) and indexed by row too (main.lua:1012, adcensus.cu:576): H <= W required")
    if out is None:
        out = np.zeros_like(vol_hwd)
    tmp = np.empty((W, D), np.float32)
    cf = ctypes.c_float
    lib().orc_sgm2(_p(x0), _p(x1), _p(vol_hwd), _p(out), _p(tmp), H, W, D, cf(pi1), cf(pi2), cf(tau_so),
                   cf(alpha1), cf(sgm_q1), cf(sgm_q2), direction)
    return out


def sgm2_band(x0, x1, vol_hwd, out, Wt, yoff, xoff, pi1, pi2, tau_so, alpha1, sgm_q1, sgm_q2, direction, pass_mask):
    """selected passes over a band volume (H,W,D) at image offset (yoff, xoff); x0/x1 are the full images"""
    x0, x1 = _f(x0), _f(x1)
    assert vol_hwd.dtype == np.float32 and vol_hwd.flags["C_CONTIGUOUS"] and out.flags["C_CONTIGUOUS"]
    H, W, D = vol_hwd.shape
    tmp = np.empty((max(H, W), D), np.float32)
    cf = ctypes.c_float
    lib().orc_sgm2_band(_p(x0), _p(x1), _p(vol_hwd), _p(out), _p(tmp), H, W, D, int(Wt), int(yoff), int(xoff),
                        cf(pi1), cf(pi2), cf(tau_so), cf(alpha1), cf(sgm_q1), cf(sgm_q2), int(direction), int(pass_mask))
    return out


def sgm2_vrows(x0, x1, vol_hwd, out, state, Ht, yoff, pi1, pi2, tau_so, alpha1, sgm_q1, sgm_q2, direction, sd, xa, xb):
    """one vertical pass (sd 2 down / 3 up) over the row band [yoff, yoff + H) of an Ht-row image, columns [xa, xb);
    `state` (D, W) float32 carries the line state in and out (the CPU statement of the GPU's chained vertical scans)"""
    x0, x1 = _f(x0), _f(x1)
    assert vol_hwd.dtype == np.float32 and vol_hwd.flags["C_CONTIGUOUS"] and out.flags["C_CONTIGUOUS"]
    assert state.dtype == np.float32 and state.flags["C_CONTIGUOUS"]
    H, W, D = vol_hwd.shape
    cf = ctypes.c_float
    lib().orc_sgm2_vrows(_p(x0), _p(x1), _p(vol_hwd), _p(out), _p(state), H, W, D, int(Ht), int(yoff), cf(pi1), cf(pi2), cf(tau_so),
                         cf(alpha1), cf(sgm_q1), cf(sgm_q2), int(direction), int(sd), int(xa), int(xb))
    return out


def spatial_argmin(vol):
    vol = _f(vol)
    D, H, W = vol.shape
    out = np.empty((H, W), np.float32)
    lib().orc_spatial_argmin(_p(vol), _p(out), D, H * W)
    return out


def outlier_detection(d0, d1, disp_max):
    d0, d1 = _f(d0), _f(d1)
    H, W = d0.shape
    out = np.empty_like(d0)
    lib().orc_outlier_detection(_p(d0), _p(d1), _p(out), H, W, disp_max)
    return out


def interpolate_occlusion(d0, outlier):
    d0, outlier = _f(d0), _f(outlier)
    H, W = d0.shape
    out = np.empty_like(d0)
    lib().orc_interpolate_occlusion(_p(d0), _p(outlier), _p(out), H, W)
    return out


def interpolate_mismatch(d0, outlier):
    d0, outlier = _f(d0), _f(outlier)
    H, W = d0.shape
    out = np.empty_like(d0)
    lib().orc_interpolate_mismatch(_p(d0), _p(outlier), _p(out), H, W)
    return out


def subpixel_enchancement(d0, vol, disp_max):
    d0, vol = _f(d0), _f(vol)
    H, W = d0.shape
    out = np.empty_like(d0)
    lib().orc_subpixel_enchancement(_p(d0), _p(vol), _p(out), H, W, disp_max)
    return out


def median2d(img, ksize):
    img = _f(img)
    H, W = img.shape
    out = np.empty_like(img)
    lib().orc_median2d(_p(img), _p(out), H, W, ksize)
    return out


def gaussian(sigma):
    l = lib()
    l.orc_gaussian.restype = ctypes.c_int
    ks = l.orc_gaussian(ctypes.c_double(sigma), None)
    out = np.empty((ks, ks), np.float32)
    l.orc_gaussian(ctypes.c_double(sigma), _p(out))
    return out


def mean2d(img, kernel, alpha2):
    img, kernel = _f(img), _f(kernel)
    H, W = img.shape
    out = np.empty_like(img)
    lib().orc_mean2d(_p(img), _p(kernel), _p(out), H, W, kernel.shape[0], ctypes.c_float(alpha2))
    return out


def transpose_dhw_to_hwd(vol):
    vol = _f(vol)
    D, H, W = vol.shape
    out = np.empty((H, W, D), np.float32)
    lib().orc_transpose_dhw_to_hwd(_p(vol), _p(out), D, H, W)
    return out


def transpose_hwd_to_dhw_div4(vol):
    vol = _f(vol)
    H, W, D = vol.shape
    out = np.empty((D, H, W), np.float32)
    lib().orc_transpose_hwd_to_dhw_div4(_p(vol), _p(out), D, H, W)
    return out


def stereo_predict(featL, featR, imgL, imgR, D, params, want_vols=False):
    """main.lua:929-1082 (arch 'fast') from the tower output.  Returns disp (H,W)
    [, volL, volR (D,H,W)]."""
    featL, featR, imgL, imgR = _f(featL), _f(featR), _f(imgL), _f(imgR)
    C, H, W = featL.shape
    disp = np.empty((H, W), np.float32)
    volL = np.empty((D, H, W), np.float32) if want_vols else None
    volR = np.empty((D, H, W), np.float32) if want_vols else None
    l = lib()
    l.orc_stereo_predict.restype = ctypes.c_int
    rc = l.orc_stereo_predict(_p(featL), _p(featR), _p(imgL), _p(imgR), C, D, H, W, ctypes.byref(params),
                              _p(disp), _p(volL) if want_vols else None, _p(volR) if want_vols else None)
    if rc != 0:
        raise ValueError("orc_stereo_predict: bad arguments")
    return (disp, volL, volR) if want_vols else disp


def stereo_predict_chain(imgL, imgR, D, params, arch="fast", featL=None, featR=None, want_vols=False, volL=None, volR=None):
    """main.lua:929-1082 composed from the per-operator oracle functions (the C `orc_stereo_predict`
    is the same chain for arch 'fast' in one call; tests/test_oracle_properties.py checks that the
    two agree).  arch 'ad' / 'census' (main.lua:932-942) build the volumes from the images.
    Returns disp (H,W) [, volL, volR (D,H,W)]."""
    imgL, imgR = _f(imgL), _f(imgR)
    H, W = imgL.shape
    if arch == "ad":
        volL, volR = ad(imgL, imgR, D, -1), ad(imgR, imgL, D, 1)                   # :934-935
    elif arch == "census":
        volL, volR = census(imgL, imgR, D, -1), census(imgR, imgL, D, 1)           # :940-941
    elif arch == "fast":
        volL, volR = stereo_join(featL, featR, D)                                  # :946-947
        fix_border(volL, params.border, -1)                                        # :948
        fix_border(volR, params.border, 1)                                         # :949
    elif arch == "volumes":                                                        # arch 'slow': the scorer head's volumes (:962-979) are given
        volL, volR = np.array(volL, np.float32, copy=True), np.array(volR, np.float32, copy=True)
        fix_border(volL, params.border, -1)                                        # :981
        fix_border(volR, params.border, 1)
    else:
        raise ValueError(arch)
    x0c, x1c = cross(imgL, params.L1, params.tau1), cross(imgR, params.L1, params.tau1)   # :995-996
    disp = {}
    vols = {}
    for direction in (1, -1):                                                      # :955
        vol = volL if direction == -1 else volR                                    # :986
        for _ in range(params.cbca_i1):                                            # :998-1001
            vol = cbca(x0c, x1c, vol, direction)
        for _ in range(params.sgm_i):                                              # :1008-1020
            out = sgm2(imgL, imgR, transpose_dhw_to_hwd(vol), params.pi1, params.pi2, params.tau_so, params.alpha1,
                       params.sgm_q1, params.sgm_q2, direction)
            vol = transpose_hwd_to_dhw_div4(out)
        for _ in range(params.cbca_i2):                                            # :1035-1038
            vol = cbca(x0c, x1c, vol, direction)
        vols[direction] = vol
        disp[direction] = spatial_argmin(vol) - 1                                  # :1049-1050
    d = disp[-1]
    if params.lr_check:                                                            # :1054-1066
        outlier = outlier_detection(disp[-1], disp[1], D)
        d = interpolate_occlusion(d, outlier)
        d = interpolate_mismatch(d, outlier)
    d = subpixel_enchancement(d, vols[-1], D)                                      # :1068
    d = median2d(d, 5)                                                             # :1073
    d = mean2d(d, gaussian(params.blur_sigma), params.blur_t)                      # :1078
    return (d, vols[-1], vols[1]) if want_vols else d
