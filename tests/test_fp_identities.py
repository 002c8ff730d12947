"""CPU: the IEEE-754 identities the bit-exactness arguments of the kernels lean on.

* CBCA accumulates with fmaf(w, q, acc), q in {0, 1}, instead of a predicated add:
  fmaf(w, 1, acc) == acc + w and fmaf(w, 0, acc) == acc (acc never -0, w finite).
* The paired horizontal SGM scans store (0 + right) + (0 + left) where the reference has
  (0 + right) + left: (0 + a) + b == (0 + b) + a.
* SGM carries padding slots as NaN: fminf ignores them like the reference's bounds guards.
"""
import ctypes
import ctypes.util

import numpy as np

_libm = ctypes.CDLL(ctypes.util.find_library("m"))
_libm.fmaf.restype = ctypes.c_float
_libm.fmaf.argtypes = [ctypes.c_float] * 3
_libm.fminf.restype = ctypes.c_float
_libm.fminf.argtypes = [ctypes.c_float] * 2


def _samples(n, seed):
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    x = bits.view(np.float32)
    x = x[np.isfinite(x)]
    special = np.array([0.0, -0.0, 1.0, -1.0, 1e-45, -1e-45, 1.17549435e-38, 3.4028235e38, -3.4028235e38,
                        0.5, 1.5, 2.0 ** 24, 2.0 ** 24 + 2, 1e-20, -1e-20], np.float32)
    near = (rng.standard_normal(n // 2) * rng.choice([1e-3, 1.0, 1e3], n // 2)).astype(np.float32)
    return np.concatenate([special, near, x])


def _bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def test_fma_with_unit_weight_is_a_plain_add():
    w, acc = _samples(4000, 0), _samples(4000, 1)
    m = min(w.size, acc.size)
    for a, b in zip(w[:m], acc[:m]):
        got = np.float32(_libm.fmaf(a, 1.0, b))
        with np.errstate(over="ignore", invalid="ignore"):
            want = np.float32(b) + np.float32(a)
        assert _bits(got) == _bits(want) or (np.isnan(got) and np.isnan(want)), (a, b)


def test_fma_with_zero_weight_keeps_a_non_negative_zero_accumulator():
    w, acc = _samples(4000, 2), _samples(4000, 3)
    m = min(w.size, acc.size)
    for a, b in zip(w[:m], acc[:m]):
        if np.signbit(b) and b == 0:
            continue                                      # -0 accumulators never occur (they start at +0)
        got = np.float32(_libm.fmaf(a, 0.0, b))
        assert _bits(got) == _bits(np.float32(b)), (a, b)
    # and why they never occur: a sum is -0 only if both operands are -0
    for a in w[:500]:
        s = np.float32(0.0) + np.float32(a)
        assert not (s == 0 and np.signbit(s)), a


def test_zero_plus_a_plus_b_commutes():
    a, b = _samples(6000, 4), _samples(6000, 5)
    m = min(a.size, b.size)
    z = np.float32(0.0)
    with np.errstate(over="ignore", invalid="ignore"):
        left = (z + a[:m]) + b[:m]                        # (0 + right) + left, the reference's order
        right = (z + a[:m]) + (z + b[:m])                 # what the scan that arrives second computes
        swapped = (z + b[:m]) + a[:m]
    for x in (right, swapped):
        same = (_bits(left) == _bits(x)) | (np.isnan(left) & np.isnan(x))
        assert same.all()


def test_fminf_ignores_nan_padding():
    vals = _samples(300, 6)
    nan = np.float32(np.nan)
    for v in vals[:300]:
        assert _bits(np.float32(_libm.fminf(v, nan))) == _bits(v) or v == 0
        assert _bits(np.float32(_libm.fminf(nan, v))) == _bits(v) or v == 0
    assert np.isnan(_libm.fminf(nan, nan))
