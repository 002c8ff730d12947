"""CPU: the numerical model behind the tensor-core kernels (csrc/scorer_head.cu, feature_tower.cu, umma.cuh): an fp32 operand x
is split into two bf16 values, x = hi + lo (hi = bf16(x), lo = bf16(x - hi)), and a product a . w is evaluated as
hi_a.hi_w + lo_a.hi_w + hi_a.lo_w with fp32 accumulation (three bf16 MMAs into one accumulator).  The dropped lo_a.lo_w term and
the rounding of lo are ~2^-16 relative, so a 384-term dot product keeps ~1e-6 relative accuracy, against ~4e-3 for plain bf16 --
which is what lets those kernels meet the north star's 1e-4 on their outputs."""
import numpy as np


def bf16(x):
    """round-to-nearest-even float32 -> bfloat16, returned as float32 (the prep kernels' and cvt.rn.bf16x2's rounding)"""
    b = np.asarray(x, np.float32).view(np.uint32)
    r = (b + np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)
    return r.view(np.float32)


def split(x):
    hi = bf16(x)
    return hi, bf16(x.astype(np.float32) - hi)


def test_split_reconstructs_to_sixteen_bits():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(100000) * 10 ** rng.uniform(-3, 3, 100000)).astype(np.float32)
    hi, lo = split(x)
    assert np.all(np.abs((hi.astype(np.float64) + lo) - x) <= np.abs(x) * 2.0 ** -16)


def test_three_term_product_is_fp32_grade():
    rng = np.random.default_rng(1)
    K, M, N = 384, 64, 48
    a = np.maximum(rng.standard_normal((M, K)), 0).astype(np.float32)            # post-ReLU activations
    w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    want = a.astype(np.float64) @ w.astype(np.float64).T
    ah, al = split(a)
    wh, wl = split(w)
    f32 = lambda p, q: (p.astype(np.float32) @ q.astype(np.float32).T)              # fp32 accumulation of exact bf16 products
    got3 = f32(ah, wh) + f32(al, wh) + f32(ah, wl)
    got1 = f32(ah, wh)
    scale = np.abs(want).max()
    assert np.abs(got3 - want).max() <= 1e-5 * scale
    assert np.abs(got1 - want).max() >= 1e-4 * scale                               # plain bf16 is NOT enough for the 1e-4 bar
