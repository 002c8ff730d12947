"""CPU: the C-ABI shared library builds, loads, and exports every symbol that
include/adcensus_b200.h declares; the host-side mirror exposes every hot-path name of the
reference's funcs[] table (adcensus.cu:2061-2096).  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

import mccnn_b200
from mccnn_b200 import adcensus, pipeline

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    mccnn_b200.build()
    return adcensus.lib()


def declared_functions():
    src = open(os.path.join(ROOT, "include", "adcensus_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:adcensus|mccnn)_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libadcensus_b200.so does not export %s" % n


def test_version(lib):
    assert "sm_100a" in adcensus.version()


HOT_PATH_LUA_NAMES = ["StereoJoin", "cross", "cbca", "sgm2", "outlier_detection", "interpolate_occlusion",
                      "interpolate_mismatch", "subpixel_enchancement", "median2d", "mean2d", "Normalize_forward",
                      "spatial_argmin", "ad", "census"]


def test_python_mirror_has_reference_names():
    for n in HOT_PATH_LUA_NAMES:
        assert callable(getattr(adcensus, n)), n


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    # null pointers / bad sizes return ADCENSUS_EINVAL before any CUDA call
    assert lib.adcensus_StereoJoin(None, None, None, None, 64, 16, 64, 128, None) == -1
    assert lib.adcensus_cross(None, None, 4, 4, 5, ctypes.c_float(0.1), None) == -1
    one = ctypes.c_void_p(16)
    assert lib.adcensus_StereoJoin(one, one, one, one, 129, 16, 64, 128, None) == -2      # C > 128
    assert lib.adcensus_sgm2(one, one, one, ctypes.c_void_p(32), None, 4, 8, 513, ctypes.c_float(1), ctypes.c_float(1),
                             ctypes.c_float(1), ctypes.c_float(1), ctypes.c_float(1), ctypes.c_float(1), 1, None) == -2
    assert lib.adcensus_median2d(one, one, 4, 4, 13, None) == -2
    assert lib.adcensus_median2d(one, one, 4, 4, 4, None) == -1                            # even kernel
    assert lib.adcensus_cbca(one, one, one, one, 4, 4, 4, 1, None) == -1                   # in aliases out


def test_round2_entry_points_reject_bad_arguments_without_a_gpu(lib):
    """scorer head / feature tower / batch call: EINVAL (-1) for null or nonsensical arguments, ELIMIT (-2) for shapes the
    kernels are not built for -- all decided before the first CUDA call"""
    h = ctypes.c_void_p()
    arr = (ctypes.c_void_p * 5)(*[16] * 5)
    lib.mccnn_scorer_head_create.restype = lib.mccnn_feature_tower_create.restype = ctypes.c_int
    sh = lambda fm, nh2, l2, W=arr, b=arr: lib.mccnn_scorer_head_create(ctypes.byref(h), fm, nh2, l2, W, b, 0, None)
    assert sh(112, 384, 3, None) == -1 and sh(112, 384, 3, arr, None) == -1
    assert lib.mccnn_scorer_head_create(None, 112, 384, 3, arr, arr, 0, None) == -1
    assert sh(112, 384, 0) == -1 and sh(4, 384, 3) == -1 and sh(112, 64, 3) == -1
    assert sh(112, 200, 3) == -2                                   # nh2 must be a multiple of 128
    assert sh(108, 384, 3) == -2                                   # fm must be a multiple of 8
    assert sh(112, 384, 9) == -2                                   # more hidden layers than the kernel unrolls
    assert h.value is None
    ft = lambda n_in, fm, l1, W=arr, b=arr: lib.mccnn_feature_tower_create(ctypes.byref(h), n_in, fm, l1, 0, 1, W, b, 0, None)
    assert ft(1, 64, 4, None) == -1 and ft(2, 64, 4) == -1 and ft(1, 64, 0) == -1 and ft(1, 4, 4) == -1
    assert ft(1, 72, 4) == -2                                      # fm must be a multiple of 16
    assert ft(1, 64, 40) == -2
    assert h.value is None
    assert lib.mccnn_scorer_head_forward(None, arr, arr, arr, arr, 4, 4, 4, 3, None) == -1
    assert lib.mccnn_feature_tower_forward(None, arr, arr, 2, 4, 4, 3, None) == -1
    assert lib.mccnn_pipeline_run_batch(None, 1, arr, arr, arr, arr, arr, None) == -1


def test_python_mirror_type_errors():
    import torch

    with pytest.raises(adcensus.AdcensusError, match="torch.CudaTensor expected"):
        adcensus.StereoJoin(torch.zeros(1, 2, 3, 4), torch.zeros(1, 2, 3, 4), torch.zeros(1, 2, 3, 4), torch.zeros(1, 2, 3, 4))


def test_presets_match_main_lua():
    p = pipeline.make_params("kitti", "fast")
    assert (p.L1, p.cbca_i1, p.cbca_i2, p.sgm_i) == (0, 0, 0, 1)
    assert abs(p.pi2 - 55.72) < 1e-5 and abs(p.blur_sigma - 7.74) < 1e-12
    p = pipeline.make_params("mb", "slow")
    assert (p.L1, p.cbca_i2, p.lr_check, p.border) == (14, 16, 0, 5)


def test_gaussian_host_helper(lib):
    from oracle import oracle as orc

    import numpy as np

    for s in (0.5, 1.67, 5.99, 7.74):
        assert np.array_equal(adcensus.gaussian(s).numpy(), orc.gaussian(s))
