"""GPU, 2 processes over NCCL (skipped on a single-GPU box): the row-band split of one pair must be
bit-identical to the single-GPU fused pipeline; the batch path must scale by sharding pairs."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need2():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def _torchrun(script_args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + script_args
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("args", [
    ["--H", "96", "--W", "160", "--D", "24", "--C", "16", "--preset", "kitti:slow", "--cbca_i2", "2"],
    ["--H", "120", "--W", "200", "--D", "40", "--C", "8", "--preset", "mb:slow", "--cbca_i2", "2"],
])
def test_rowband_two_gpus_bit_identical(args):
    _need2()
    res = _torchrun(["tools/run_rowband.py", "--check", "--iters", "1"] + args, 29561)
    assert res["n_gpus"] == 2 and res["mismatches_vs_single_gpu"] == 0 and res["finite"]


def test_bench_two_gpus_prints_one_json_line():
    _need2()
    res = _torchrun(["bench.py", "--gpus", "2", "--steps", "2", "--warmup", "3", "--small", "--no-cpu-baseline"], 29562)
    assert res["n_gpus"] == 2 and res["scaling"] == "weak" and res["value"] > 0 and res["gpu_launches"] > 0
