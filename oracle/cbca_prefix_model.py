"""Numerical model of an O(1)-per-pixel CBCA (prefix sums along x, then along y) -- TEST INFRASTRUCTURE
ONLY; a design study for a tolerance-mode kernel, no product code uses it.

cbca (adcensus.cu:343-377) sums vol[d] over a cross-shaped support: rows yy in (yy_s, yy_t), and in every
row the run xx in (xx_s(yy), xx_t(yy)).  The run of row yy depends on (d, yy, x) only, NOT on which output
row y uses it, so with
    S(d, yy, x) = sum of vol[d, yy, xx_s+1 .. xx_t-1]   = P(yy, xx_t) - P(yy, xx_s + 1)      (P: prefix along x)
    C(d, yy, x) = number of taps of that run
the output is a difference of prefix sums of S along y:
    out(d, y, x) = (T(yy_t, x) - T(yy_s + 1, x)) / (N(yy_t, x) - N(yy_s + 1, x))             (T, N: prefix of S, C along y)
i.e. a constant amount of work per pixel instead of one addition per tap -- but a different summation
order, so it is only within rounding of the reference (the north star allows 1e-4 relative for float
aggregation; index outputs stay bit-exact only as far as no arg-min is decided by that rounding).
`tile` limits the length of every prefix (a kernel works on tiles with a halo), which bounds the
cancellation error; `dtype` is the precision the prefixes are held in.
"""
import numpy as np


def cbca_prefix(x0c, x1c, vol, direction, dtype=np.float32, tile=(32, 128)):
    x0c, x1c, vol = (np.asarray(a, np.float32) for a in (x0c, x1c, vol))
    D, H, W = vol.shape
    out = vol.copy()                                                       # x' outside the image: plain copy (:353-354)
    ys, xs = np.mgrid[0:H, 0:W]
    for d in range(D):
        sh = d * direction
        xp = xs + sh
        valid = (xp >= 0) & (xp < W)
        xpc = np.clip(xp, 0, W - 1)
        # per (yy, x): the run of row yy (:362-363) and the vertical range of output (y, x) (:359-360)
        xx_s = np.maximum(x0c[0], x1c[0][ys, xpc] - sh).astype(np.int64)
        xx_t = np.minimum(x0c[1], x1c[1][ys, xpc] - sh).astype(np.int64)
        yy_s = np.maximum(x0c[2], x1c[2][ys, xpc]).astype(np.int64)
        yy_t = np.minimum(x0c[3], x1c[3][ys, xpc]).astype(np.int64)
        v = np.where(np.isnan(vol[d]), np.float32(0), vol[d])
        ty, tx = tile
        res = np.full((H, W), np.nan, np.float32)
        for y0 in range(0, H, ty):
            for x0 in range(0, W, tx):
                y1, x1 = min(y0 + ty, H), min(x0 + tx, W)
                # halo: everything the outputs of this tile can reach
                ya, yb = max(0, int(yy_s[y0:y1, x0:x1].min()) + 1), min(H, int(yy_t[y0:y1, x0:x1].max()))
                if yb <= ya:
                    ya, yb = y0, y1
                xa = max(0, int(xx_s[ya:yb, x0:x1].min()) + 1)
                xb = min(W, int(xx_t[ya:yb, x0:x1].max()))
                if xb <= xa:
                    xa, xb = x0, x1
                P = np.zeros((yb - ya, xb - xa + 1), dtype)
                P[:, 1:] = np.cumsum(v[ya:yb, xa:xb].astype(dtype), axis=1, dtype=dtype)
                rows = np.arange(ya, yb)[:, None] - ya
                lo = np.clip(xx_s[ya:yb, x0:x1] + 1, xa, xb) - xa
                hi = np.clip(xx_t[ya:yb, x0:x1], xa, xb) - xa
                hi = np.maximum(hi, lo)
                S = (P[rows, hi] - P[rows, lo]).astype(dtype)               # row sums of the runs, every row of the halo
                C = (hi - lo).astype(np.int64)
                T = np.zeros((yb - ya + 1, x1 - x0), dtype)
                T[1:] = np.cumsum(S, axis=0, dtype=dtype)
                N = np.zeros((yb - ya + 1, x1 - x0), np.int64)
                N[1:] = np.cumsum(C, axis=0)
                cols = np.arange(x1 - x0)[None, :]
                a = np.clip(yy_s[y0:y1, x0:x1] + 1, ya, yb) - ya
                b = np.clip(yy_t[y0:y1, x0:x1], ya, yb) - ya
                b = np.maximum(b, a)
                num = (T[b, cols] - T[a, cols]).astype(np.float32)
                den = (N[b, cols] - N[a, cols]).astype(np.float32)
                with np.errstate(invalid="ignore", divide="ignore"):
                    res[y0:y1, x0:x1] = num / den
        out[d] = np.where(valid, res, vol[d])
    return out
