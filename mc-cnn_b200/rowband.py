"""One stereo pair split by ROW BANDS over N GPUs (BASELINE.json config 5: a Middlebury-scale
volume), one process per GPU, torch.distributed (NCCL over NVLink) for the exchanges.

Which stages need an exchange follows from the reference's data dependences, not
from BASELINE.json's wording:

* StereoJoin, fix_border, the permutes, argmin, sub-pixel: row-local -> no communication;
* cross: needs image rows +- L1 -> both (small) images are replicated on every rank;
* cbca: the vertical arm reaches (max(L1,2)-1) rows into the neighbouring bands -> a halo of that
  many volume rows is exchanged with the two neighbours before EVERY iteration (send/recv);
* sgm2: the two horizontal passes are band-local; the two vertical passes are a serial chain over
  the rows, so the accumulator and the cost volume are RE-PARTITIONED into column bands (one
  all-to-all each, every rank sends (N-1)/N of its band), the vertical passes run on whole
  columns, and the accumulator comes back (one more all-to-all).  Direction order and therefore
  the accumulation order (right, left, down, up) are the reference's => results are bit-identical
  to the single-GPU pipeline (tests/test_rowband_*.py);
* LR check / interpolations / median / bilateral work on H x W maps: the disparity maps are
  all-gathered (tiny) and these stages run replicated; sub-pixel refinement reads the band's own
  left volume.

The operators are injected (``ops``), so the same driver runs on the CUDA library (CudaOps) and,
in the CPU tests, on any object with the same methods over CPU tensors.
"""
import torch
import torch.distributed as dist


def split(n, parts, i):
    """[lo, hi) of part i when n items are cut into `parts` nearly equal contiguous pieces"""
    return (n * i) // parts, (n * (i + 1)) // parts


class Comm:
    """Thin layer over torch.distributed point-to-point ops (works with nccl and gloo)."""

    def __init__(self):
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.rank = dist.get_rank() if self.on else 0
        self.world = dist.get_world_size() if self.on else 1

    def exchange(self, sends, recvs):
        """sends: {peer: tensor}, recvs: {peer: empty tensor to fill}; all in flight together."""
        if not self.on:
            return
        ops = [dist.P2POp(dist.isend, t.contiguous(), p) for p, t in sends.items() if p != self.rank]
        ops += [dist.P2POp(dist.irecv, t, p) for p, t in recvs.items() if p != self.rank]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def all_gather_rows(self, band, H):
        """band: this rank's rows (Hb, W) of an H x W map -> the full map on every rank"""
        if not self.on:
            return band
        W = band.shape[-1]
        full = band.new_empty((H, W))
        y0, y1 = split(H, self.world, self.rank)
        full[y0:y1] = band
        sends = {p: band for p in range(self.world)}
        recvs = {}
        for p in range(self.world):
            a, b = split(H, self.world, p)
            recvs[p] = full[a:b]
        self.exchange(sends, recvs)
        return full

    def rows_to_cols(self, t, H, W):
        """(Hb, W, D) row band -> (H, Wb, D) column band (all-to-all)"""
        if not self.on:
            return t
        D = t.shape[2]
        x0, x1 = split(W, self.world, self.rank)
        out = t.new_empty((H, x1 - x0, D))
        sends, recvs = {}, {}
        for p in range(self.world):
            a, b = split(W, self.world, p)
            sends[p] = t[:, a:b, :]
            ya, yb = split(H, self.world, p)
            recvs[p] = out[ya:yb]                      # contiguous: rows are the outermost axis
        y0, y1 = split(H, self.world, self.rank)
        out[y0:y1] = t[:, x0:x1, :]
        self.exchange(sends, recvs)
        return out

    def cols_to_rows(self, t, H, W):
        """(H, Wb, D) column band -> (Hb, W, D) row band (all-to-all back)"""
        if not self.on:
            return t
        D = t.shape[2]
        y0, y1 = split(H, self.world, self.rank)
        out = t.new_empty((y1 - y0, W, D))
        sends, recvs, stage = {}, {}, {}
        for p in range(self.world):
            ya, yb = split(H, self.world, p)
            sends[p] = t[ya:yb]
            a, b = split(W, self.world, p)
            stage[p] = t.new_empty((y1 - y0, b - a, D))  # strided destination: receive, then place
            recvs[p] = stage[p]
        self.exchange(sends, recvs)
        for p in range(self.world):
            a, b = split(W, self.world, p)
            out[:, a:b, :] = t[y0:y1] if p == self.rank else stage[p]
        return out

    def halo(self, vol, n):
        """vol (D, Hb, W): returns (rows from the band above, rows from the band below), each
        (D, <=n, W), empty at the image border"""
        D, Hb, W = vol.shape
        top = vol.new_empty((D, n if self.rank > 0 else 0, W))
        bot = vol.new_empty((D, n if self.rank < self.world - 1 else 0, W))
        if not self.on or n == 0:
            return vol.new_empty((D, 0, W)), vol.new_empty((D, 0, W))
        sends, recvs = {}, {}
        if self.rank > 0:
            sends[self.rank - 1] = vol[:, :n, :]
            recvs[self.rank - 1] = top
        if self.rank < self.world - 1:
            sends[self.rank + 1] = vol[:, Hb - n:, :]
            recvs[self.rank + 1] = bot
        # the two neighbours are distinct peers, so one batch carries both directions
        self.exchange(sends, recvs)
        return top, bot


def stereo_predict_rowband(ops, featL, featR, imgL, imgR, D, opt, comm=None):
    """main.lua:929-1082 (arch 'fast') with the rows of one pair split over the ranks of `comm`.

    featL/featR (C,H,W) and imgL/imgR (H,W) are the FULL tensors on every rank (only this rank's
    rows of the features are read).  Returns the full (H,W) disparity map on every rank.
    """
    comm = comm or Comm()
    H, W = imgL.shape
    N, r = comm.world, comm.rank
    y0, y1 = split(H, N, r)
    Hb = y1 - y0
    max_arm = max(int(opt.L1), 2)
    halo = max_arm - 1 if (opt.cbca_i1 + opt.cbca_i2) > 0 else 0
    if N > 1:
        assert min(split(H, N, p)[1] - split(H, N, p)[0] for p in range(N)) >= max(halo, 1), "bands thinner than the CBCA halo"
        assert max_arm <= 14, "row-band CBCA supports arms up to 14 pixels"

    volL, volR = ops.stereo_join(featL[:, y0:y1].contiguous(), featR[:, y0:y1].contiguous(), D)     # :946-947
    ops.fix_border(volL, opt.border, -1)                                                             # :948
    ops.fix_border(volR, opt.border, 1)                                                              # :949
    x0c = ops.cross(imgL, opt.L1, opt.tau1)                                                          # :995 (full image)
    x1c = ops.cross(imgR, opt.L1, opt.tau1)                                                          # :996

    def band_arms(arms, ya, yb):
        """arms of rows [ya, yb) in the coordinates of that sub-image; vertical end-points clamped to it
        (only halo rows are affected, whose outputs are discarded)"""
        a = arms[:, ya:yb].clone()
        a[2] = torch.clamp(a[2] - ya, min=-1)
        a[3] = torch.clamp(a[3] - ya, max=yb - ya)
        return a.contiguous()

    def cbca_band(vol, direction):
        top, bot = comm.halo(vol, halo)
        ext = torch.cat([top, vol, bot], dim=1).contiguous() if (top.shape[1] or bot.shape[1]) else vol
        ya, yb = y0 - top.shape[1], y1 + bot.shape[1]
        out = ops.cbca(band_arms(x0c, ya, yb), band_arms(x1c, ya, yb), ext, direction, max_arm)
        return out[:, top.shape[1]: top.shape[1] + Hb].contiguous()

    disp = {}
    vol_left = None
    x0b, x1b = split(W, N, r)
    for direction in (1, -1):                                                                        # :955
        vol = volL if direction == -1 else volR                                                      # :986
        for _ in range(opt.cbca_i1):                                                                 # :998-1001
            vol = cbca_band(vol, direction)
        for _ in range(opt.sgm_i):                                                                   # :1008-1020
            cost = ops.to_hwd(vol)                                                                   # (Hb, W, D)
            acc = ops.zeros_like(cost)                                                               # :1014
            ops.sgm_band(imgL, imgR, cost, acc, H, W, y0, 0, opt, direction, 3, True)                # right, left
            cost_c = comm.rows_to_cols(cost, H, W)
            acc_c = comm.rows_to_cols(acc, H, W)
            ops.sgm_band(imgL, imgR, cost_c, acc_c, H, W, 0, x0b, opt, direction, 12, False)         # down, up
            acc = comm.cols_to_rows(acc_c, H, W)
            vol = ops.from_hwd_div4(acc)                                                             # :1017-1020
        for _ in range(opt.cbca_i2):                                                                 # :1035-1038
            vol = cbca_band(vol, direction)
        disp[direction] = comm.all_gather_rows(ops.argmin(vol), H)                                   # :1049-1050
        if direction == -1:
            vol_left = vol

    d = disp[-1]
    if opt.lr_check:                                                                                 # :1054-1066
        outlier = ops.outlier_detection(disp[-1], disp[1], D)
        d = ops.interpolate_occlusion(d, outlier)
        d = ops.interpolate_mismatch(d, outlier)
    sub = ops.subpixel(d[y0:y1].contiguous(), vol_left, D)                                           # :1068 (band-local)
    d = comm.all_gather_rows(sub, H)
    d = ops.median2d(d, 5)                                                                           # :1073
    return ops.mean2d(d, opt.blur_sigma, opt.blur_t)                                                 # :1078


class CudaOps:
    """The operators on the CUDA library (torch CUDA tensors in, out)."""

    def __init__(self, device):
        self.device = device

    def stereo_join(self, fL, fR, D):
        from . import adcensus

        C, Hb, W = fL.shape
        vols = torch.empty((2, D, Hb, W), device=fL.device, dtype=torch.float32)
        adcensus.fill_nan(vols)
        adcensus.StereoJoin(fL[None], fR[None], vols[0:1], vols[1:2])
        return vols[0], vols[1]

    def fix_border(self, vol, n, direction):
        from . import adcensus

        adcensus.fix_border(vol[None], n, direction)

    def cross(self, img, L1, tau1):
        from . import adcensus

        out = torch.empty((1, 4) + tuple(img.shape), device=img.device, dtype=torch.float32)
        adcensus.cross(img[None].contiguous(), out, L1, tau1)
        return out[0]

    def cbca(self, x0c, x1c, vol, direction, max_arm):
        from . import adcensus

        out = torch.empty_like(vol)
        adcensus.cbca(x0c[None], x1c[None], vol[None], out[None], direction, max_arm=max_arm)
        return out

    def to_hwd(self, vol):
        from . import adcensus

        return adcensus.transpose_dhw_to_hwd(vol[None])[0]

    def from_hwd_div4(self, acc):
        from . import adcensus

        return adcensus.transpose_hwd_to_dhw_div4(acc[None])[0]

    def zeros_like(self, t):
        return torch.zeros_like(t)

    def sgm_band(self, imgL, imgR, cost, acc, Ht, Wt, yoff, xoff, opt, direction, pass_mask, zero_out):
        import ctypes

        from . import adcensus

        H, W, D = cost.shape
        f = ctypes.c_float
        p = lambda t: adcensus._t(t, 0, "mccnn_sgm2_band")
        with torch.cuda.device(cost.device):
            rc = adcensus.lib().mccnn_sgm2_band(p(imgL), p(imgR), p(cost), p(acc), H, W, D, Ht, Wt, yoff, xoff,
                                                f(opt.pi1), f(opt.pi2), f(opt.tau_so), f(opt.alpha1), f(opt.sgm_q1),
                                                f(opt.sgm_q2), int(direction), int(pass_mask), int(bool(zero_out)),
                                                adcensus._stream(cost))
        adcensus._check(rc, "mccnn_sgm2_band")

    def argmin(self, vol):
        from . import adcensus

        return adcensus.argmin(vol[None])[0, 0]

    def outlier_detection(self, dL, dR, D):
        from . import adcensus

        out = torch.zeros_like(dL)
        adcensus.outlier_detection(dL[None, None], dR[None, None], out[None, None], D)
        return out

    def interpolate_occlusion(self, d, outlier):
        from . import adcensus

        return adcensus.interpolate_occlusion(d[None, None].contiguous(), outlier[None, None].contiguous())[0, 0]

    def interpolate_mismatch(self, d, outlier):
        from . import adcensus

        return adcensus.interpolate_mismatch(d[None, None].contiguous(), outlier[None, None].contiguous())[0, 0]

    def subpixel(self, d_band, vol_band, D):
        from . import adcensus

        return adcensus.subpixel_enchancement(d_band[None, None].contiguous(), vol_band[None].contiguous(), D)[0, 0]

    def median2d(self, d, k):
        from . import adcensus

        return adcensus.median2d(d[None, None].contiguous(), k)[0, 0]

    def mean2d(self, d, sigma, t):
        from . import adcensus

        return adcensus.mean2d(d[None, None].contiguous(), adcensus.gaussian(sigma).to(d.device), t)[0, 0]
