"""One stereo pair split by row bands over the ranks of a torchrun job (BASELINE.json config 5).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/run_rowband.py --H 2000 --W 3000 --D 400 --preset mb:fast [--check] [--iters 3]

Every rank builds the same seeded synthetic pair, runs mccnn_b200.rowband.stereo_predict_rowband on
the CUDA library, times it with CUDA events (barrier on both sides, max over ranks) and, with
--check, rank 0 also runs the single-GPU fused pipeline and requires a bit-identical disparity map.
Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mccnn_b200  # noqa: E402,F401
from mccnn_b200 import pipeline, rowband, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--H", type=int, default=2000)
ap.add_argument("--W", type=int, default=3000)
ap.add_argument("--D", type=int, default=400)
ap.add_argument("--C", type=int, default=64)
ap.add_argument("--preset", default="mb:fast")
ap.add_argument("--L1", type=int, default=None)
ap.add_argument("--tau1", type=float, default=None)
ap.add_argument("--cbca_i1", type=int, default=None)
ap.add_argument("--cbca_i2", type=int, default=None)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--check", action="store_true")
ap.add_argument("--chunks", type=int, default=None, help="column chunks of the vertical SGM wavefront")
ap.add_argument("--cheap-inputs", action="store_true", help="random unit-norm features generated on the device (bench sizes)")
a = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
json_fd = os.dup(1)
os.dup2(2, 1)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

over = {k: getattr(a, k) for k in ("L1", "tau1", "cbca_i1", "cbca_i2") if getattr(a, k) is not None}
opt = pipeline.make_params(*a.preset.split(":"), **over)
H, W, D, C = a.H, a.W, a.D, a.C
if a.cheap_inputs:
    g = torch.Generator(device=dev).manual_seed(7)
    featL = torch.nn.functional.normalize(torch.randn((C, H, W), device=dev, generator=g), dim=0)
    featR = torch.nn.functional.normalize(torch.randn((C, H, W), device=dev, generator=g), dim=0)
    rng = np.random.default_rng(7)
    img = synth.natural_image(rng, H, W + 16)
    st = lambda x: torch.from_numpy(((x - x.mean()) / x.std(ddof=1)).astype(np.float32)).to(dev)
    imgL, imgR = st(img[:, 16:]).contiguous(), st(img[:, :W]).contiguous()
else:
    p = synth.make_pair(H, W, C, D, seed=11)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    featL, featR, imgL, imgR = t(p["featL"]), t(p["featR"]), t(p["imgL"]), t(p["imgR"])

ops = rowband.CudaOps(dev)
comm = rowband.Comm()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


out = rowband.stereo_predict_rowband(ops, featL, featR, imgL, imgR, D, opt, comm, chunks=a.chunks)   # warm-up (also the result we check)
barrier()
times = []
for _ in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    out = rowband.stereo_predict_rowband(ops, featL, featR, imgL, imgR, D, opt, comm, chunks=a.chunks)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    times.append(ms)

res = {"workload": "rowband %dx%d d=%d C=%d %s %s" % (H, W, D, C, a.preset, over), "n_gpus": world,
       "ms_min": round(min(times), 3), "ms_all": [round(x, 3) for x in times], "finite": bool(torch.isfinite(out).all())}
if a.check and rank == 0:
    sp = pipeline.StereoPipeline(C, D, H, W, opt, device=local, cbca_mode="exact")   # the band driver runs the exact operators
    ref = sp.run(featL, featR, imgL, imgR)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ref = sp.run(featL, featR, imgL, imgR)
    e1.record()
    torch.cuda.synchronize()
    res["single_gpu_fused_ms"] = round(e0.elapsed_time(e1), 3)
    bad = ~((out == ref) | (torch.isnan(out) & torch.isnan(ref)))
    res["mismatches_vs_single_gpu"] = int(bad.sum())
    sp.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
if rank == 0:
    os.write(json_fd, (json.dumps(res) + "\n").encode())
    if res.get("mismatches_vs_single_gpu", 0) != 0:
        sys.exit(1)
