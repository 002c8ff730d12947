"""GPU: the two-lane batch calls (mccnn_pipeline_run_batch / _run_host_batch) give exactly the disparity maps of one
mccnn_pipeline_run per pair."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mccnn_b200  # noqa: E402,F401
from mccnn_b200 import pipeline, synth  # noqa: E402


def test_batch_lanes_equal_single_runs():
    H, W, C, D = 60, 200, 16, 40
    opt = pipeline.make_params("kitti", "accurate_cbca4")
    dev = torch.device("cuda:0")
    pairs = []
    for seed in range(5):
        p = synth.make_pair(H, W, C, D, seed=40 + seed)
        pairs.append(tuple(torch.from_numpy(p[k]).to(dev) for k in ("featL", "featR", "imgL", "imgR")))
    sp = pipeline.StereoPipeline(C, D, H, W, opt)
    want = [sp.run(*pr).clone() for pr in pairs]
    torch.cuda.synchronize()
    got = sp.run_batch(pairs)
    torch.cuda.synchronize()
    for g, w in zip(got, want):
        assert torch.equal(torch.nan_to_num(g, nan=-1.0), torch.nan_to_num(w, nan=-1.0))
    host = [tuple(t.cpu().pin_memory() for t in pr) for pr in pairs]
    got_h = sp.run_host_batch(host)
    for g, w in zip(got_h, want):
        assert np.array_equal(np.nan_to_num(g.numpy(), nan=-1.0), np.nan_to_num(w.cpu().numpy(), nan=-1.0))
    sp.close()
