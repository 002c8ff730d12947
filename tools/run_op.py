"""Run one hot-path operator a few times at the bench workload (for ncu captures).

    ncu --set full ... python tools/run_op.py cbca [--iters 3] [--small]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mccnn_b200  # noqa: E402,F401
from mccnn_b200 import adcensus, pipeline, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("op")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--D", type=int, default=228)
ap.add_argument("--H", type=int, default=370)
ap.add_argument("--W", type=int, default=1226)
ap.add_argument("--C", type=int, default=64)
a = ap.parse_args()
H, W, D, C = a.H, a.W, a.D, a.C
dev = torch.device("cuda:0")
opt = pipeline.make_params("kitti", "accurate_cbca4")
rng = np.random.default_rng(0)
img = synth.natural_image(rng, H, W + 8)
st = lambda x: torch.from_numpy(((x - x.mean()) / x.std(ddof=1)).astype(np.float32)).to(dev)
iL, iR = st(img[:, 8:])[None, None].contiguous(), st(img[:, :W])[None, None].contiguous()
g = torch.Generator(device=dev).manual_seed(0)
fL = torch.nn.functional.normalize(torch.randn((1, C, H, W), device=dev, generator=g), dim=1)
fR = torch.nn.functional.normalize(torch.randn((1, C, H, W), device=dev, generator=g), dim=1)
vols = torch.empty((2, D, H, W), device=dev)
adcensus.fill_nan(vols)
adcensus.StereoJoin(fL, fR, vols[0:1], vols[1:2])
x0c = torch.empty((1, 4, H, W), device=dev)
x1c = torch.empty((1, 4, H, W), device=dev)
adcensus.cross(iL, x0c, opt.L1, opt.tau1)
adcensus.cross(iR, x1c, opt.L1, opt.tau1)
tmp = torch.empty((1, D, H, W), device=dev)
torch.cuda.synchronize()
for _ in range(a.iters):
    if a.op == "StereoJoin":
        adcensus.StereoJoin(fL, fR, vols[0:1], vols[1:2])
    elif a.op == "cbca":
        adcensus.cbca(x0c, x1c, vols[0:1], tmp, -1)
    elif a.op == "sgm2":
        volt = adcensus.transpose_dhw_to_hwd(vols[0:1])
        out = torch.zeros_like(volt)
        adcensus.sgm2(iL, iR, volt, out, None, opt.pi1, opt.pi2, opt.tau_so, opt.alpha1, opt.sgm_q1, opt.sgm_q2, -1)
    elif a.op == "pipeline":
        sp = pipeline.StereoPipeline(C, D, H, W, opt)
        sp.run(fL[0], fR[0], iL[0, 0], iR[0, 0])
        sp.close()
    else:
        raise SystemExit("unknown op " + a.op)
torch.cuda.synchronize()
print("done", a.op)
