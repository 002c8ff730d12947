"""CPU, world_size 2 and 3 over gloo: the row-band split of ONE stereo pair (CBCA blocks on bands extended by
iterations x halo rows, one halo exchange per block; the vertical SGM passes as a wavefront over column chunks
that hands only the line state across a band boundary; gathered post-processing) with the CPU oracle standing in
for the CUDA operators.  The result must equal the
single-process oracle pipeline bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleOps:
    """rowband's operator interface on the CPU oracle (torch CPU tensors in and out)."""

    def __init__(self, orc):
        self.o = orc

    @staticmethod
    def _n(t):
        return np.ascontiguousarray(t.numpy(), dtype=np.float32)

    def stereo_join(self, fL, fR, D):
        a, b = self.o.stereo_join(self._n(fL), self._n(fR), D)
        return torch.from_numpy(a), torch.from_numpy(b)

    def fix_border(self, vol, n, direction):
        v = self._n(vol)
        self.o.fix_border(v, n, direction)
        vol.copy_(torch.from_numpy(v))

    def cross(self, img, L1, tau1):
        return torch.from_numpy(self.o.cross(self._n(img), L1, tau1))

    def cbca(self, x0c, x1c, vol, direction, max_arm):
        return torch.from_numpy(self.o.cbca(self._n(x0c), self._n(x1c), self._n(vol), direction))

    def to_hwd(self, vol):
        return torch.from_numpy(self.o.transpose_dhw_to_hwd(self._n(vol)))

    def from_hwd_div4(self, acc):
        return torch.from_numpy(self.o.transpose_hwd_to_dhw_div4(self._n(acc)))

    def zeros_like(self, t):
        return torch.zeros_like(t)

    def sgm_band(self, imgL, imgR, cost, acc, Ht, Wt, yoff, xoff, opt, direction, pass_mask, zero_out):
        c, a = self._n(cost), self._n(acc)
        if zero_out:
            a[:] = 0
        self.o.sgm2_band(self._n(imgL), self._n(imgR), c, a, Wt, yoff, xoff, opt.pi1, opt.pi2, opt.tau_so, opt.alpha1,
                         opt.sgm_q1, opt.sgm_q2, direction, pass_mask)
        acc.copy_(torch.from_numpy(a))

    def sgm_tables(self, imgL, imgR, D, opt, direction):
        return (self._n(imgL), self._n(imgR))          # the oracle looks the images up directly

    def new_state(self, W, D, like):
        return torch.zeros((W, D), dtype=torch.float32)

    def sgm_rows(self, tab, cost, acc, Ht, yoff, opt, direction, pass_mask, zero_out, xa, xb, state_in, state_out):
        iL, iR = tab
        c, a = self._n(cost), self._n(acc)
        W = c.shape[1]
        if zero_out:
            a[:] = 0
        args = (opt.pi1, opt.pi2, opt.tau_so, opt.alpha1, opt.sgm_q1, opt.sgm_q2, direction)
        if pass_mask & 3:
            self.o.sgm2_band(iL, iR, c, a, W, yoff, 0, *args, pass_mask & 3)
        for sd in (2, 3):
            if not pass_mask & (1 << sd):
                continue
            st = np.zeros((c.shape[2], W), np.float32)     # the oracle's line state is (D, W)
            if state_in is not None:
                st[:, xa:xb] = state_in[xa:xb].numpy().T
            self.o.sgm2_vrows(iL, iR, c, a, st, Ht, yoff, *args, sd, xa, xb)
            if state_out is not None:
                state_out[xa:xb] = torch.from_numpy(np.ascontiguousarray(st[:, xa:xb].T))
        acc.copy_(torch.from_numpy(a))

    def argmin(self, vol):
        return torch.from_numpy(self.o.spatial_argmin(self._n(vol)) - 1)

    def outlier_detection(self, dL, dR, D):
        return torch.from_numpy(self.o.outlier_detection(self._n(dL), self._n(dR), D))

    def interpolate_occlusion(self, d, outlier):
        return torch.from_numpy(self.o.interpolate_occlusion(self._n(d), self._n(outlier)))

    def interpolate_mismatch(self, d, outlier):
        return torch.from_numpy(self.o.interpolate_mismatch(self._n(d), self._n(outlier)))

    def subpixel(self, d_band, vol_band, D):
        return torch.from_numpy(self.o.subpixel_enchancement(self._n(d_band), self._n(vol_band), D))

    def median2d(self, d, k):
        return torch.from_numpy(self.o.median2d(self._n(d), k))

    def mean2d(self, d, sigma, t):
        return torch.from_numpy(self.o.mean2d(self._n(d), self.o.gaussian(sigma), t))


CASES = {
    "kitti_slow": (("kitti", "slow"), dict(cbca_i2=2), 37, 52, 6, 14),
    "mb_slow": (("mb", "slow"), dict(cbca_i2=2, L1=6), 41, 48, 4, 9),
    "kitti_fast": (("kitti", "fast"), {}, 23, 40, 4, 10),
}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(case):
    import mccnn_b200  # noqa: F401
    from mccnn_b200 import pipeline, synth

    preset, over, H, W, C, D = CASES[case]
    opt = pipeline.make_params(*preset, **over)
    p = synth.make_pair(H, W, C, D, seed=H + D)
    return opt, p, D


def _worker(rank, world, port, case, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mccnn_b200 import rowband
    from oracle import oracle as orc

    opt, p, D = _problem(case)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    out = rowband.stereo_predict_rowband(OracleOps(orc), t(p["featL"]), t(p["featR"]), t(p["imgL"]), t(p["imgR"]), D, opt)
    np.save(os.path.join(outdir, "disp_rank%d.npy" % rank), out.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,case", [(2, "kitti_slow"), (3, "mb_slow"), (2, "kitti_fast")])
def test_rowband_matches_single_process(tmp_path, world, case):
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc

    port = _free_port()
    mp.spawn(_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    opt, p, D = _problem(case)
    want = orc.stereo_predict(p["featL"], p["featR"], p["imgL"], p["imgR"], D, orc.Params(**opt.as_dict()))
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "disp_rank%d.npy" % r))
        assert np.array_equal(got, want, equal_nan=True), "rank %d of %d differs from the single-process pipeline" % (r, world)


def test_single_rank_is_the_plain_pipeline():
    """world 1 (no process group): the band driver degenerates to the ordinary chain"""
    sys.path.insert(0, ROOT)
    from mccnn_b200 import rowband
    from oracle import oracle as orc

    opt, p, D = _problem("kitti_slow")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    got = rowband.stereo_predict_rowband(OracleOps(orc), t(p["featL"]), t(p["featR"]), t(p["imgL"]), t(p["imgR"]), D, opt)
    want = orc.stereo_predict(p["featL"], p["featR"], p["imgL"], p["imgR"], D, orc.Params(**opt.as_dict()))
    assert np.array_equal(got.numpy(), want, equal_nan=True)


def test_split_partitions():
    from mccnn_b200 import rowband

    for n in (1, 7, 370, 2000):
        for parts in (1, 2, 3, 8):
            if parts > n:
                continue
            edges = [rowband.split(n, parts, i) for i in range(parts)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(parts - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
