"""BASELINE.json config 4: a batch of K228 stereo pairs sharded over the GPUs of one node, one process per
GPU, no collective on the data path (mc-cnn_b200/batch.py); prints one JSON line with whole-job pairs/s.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_batch.py --pairs 194            # or: python tools/run_batch.py --pairs 24   (one GPU)

Every rank synthesises its own pairs (seeds 1000 + index, so the set does not depend on the world size), keeps
two distinct pairs resident in pinned host memory and pushes its share through the host-buffer batch call
(H2D of features and images and D2H of the disparity map inside the timed region, copies overlapped with the
kernels).  Timed with a barrier + synchronize on both sides, max over ranks.
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
from mccnn_b200 import batch, pipeline, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=194)
    ap.add_argument("--D", type=int, default=228)
    ap.add_argument("--H", type=int, default=370)
    ap.add_argument("--W", type=int, default=1226)
    ap.add_argument("--C", type=int, default=64)
    a = ap.parse_args()
    json_fd = os.dup(1)
    os.dup2(2, 1)                                           # keep stdout clean for the JSON line
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    opt = pipeline.make_params("kitti", "accurate_cbca4")
    mine = batch.shard_indices(a.pairs, rank, world)
    # two distinct resident pairs per rank stand in for its whole share (the data path is identical per pair)
    host = []
    for i in mine[:2] or [rank]:
        p = synth.make_pair(a.H, a.W, a.C, a.D, seed=1000 + i)
        host.append(tuple(torch.from_numpy(np.ascontiguousarray(p[k])).pin_memory() for k in ("featL", "featR", "imgL", "imgR")))
    work = [host[j % len(host)] for j in range(len(mine))]
    disps = [torch.empty((a.H, a.W), dtype=torch.float32).pin_memory() for _ in mine]
    sp = pipeline.StereoPipeline(a.C, a.D, a.H, a.W, opt, device=local)
    sp.run_host_batch(work[:2], disps[:2])                  # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if work:
        sp.run_host_batch(work, disps)
    torch.cuda.synchronize()
    dt = batch.max_over_ranks(time.perf_counter() - t0, device="cuda")
    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {"metric": "stereo pairs/sec (370x1226 d=228), batch of %d pairs" % a.pairs, "value": round(a.pairs / dt, 3),
                "unit": "pairs/s", "n_gpus": world, "seconds": round(dt, 4), "pairs": a.pairs,
                "config": {"workload": "BASELINE config 4", "H": a.H, "W": a.W, "D": a.D, "C": a.C,
                           "parallelism": "pairs round-robin over GPUs, no collective", "path": "mccnn_pipeline_run_host_batch"}}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    sp.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
