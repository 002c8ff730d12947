"""Generate tests/golden/*.npz by running THE REFERENCE ITSELF on a B200.

TEST INFRASTRUCTURE.  Run on the GPU box (the reference's kernels are CUDA-only):

    gpurun -- python oracle/make_golden.py gpurun_out/golden

It drives oracle/_ref/libadcensus_ref.so -- /root/reference/adcensus.cu compiled
unmodified against oracle/refshim (recipe: oracle/Makefile, target `ref`) -- through
oracle/refdriver.py on small seeded inputs and stores inputs + every intermediate
tensor of main.lua's stereo_predict chain, plus a few stand-alone operator cases.
The .npz files are then committed under tests/golden/ and pin the CPU oracle
(tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_parity.py).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mccnn_b200  # noqa: E402,F401
from mccnn_b200 import pipeline, synth  # noqa: E402
from oracle import refdriver  # noqa: E402

CASES = [
    # name, H, W, C, D, preset, overrides, seed
    ("pipe_kitti_slow", 24, 40, 8, 12, ("kitti", "slow"), dict(cbca_i2=1), 11),
    ("pipe_kitti_fast", 20, 36, 6, 10, ("kitti", "fast"), {}, 12),
    ("pipe_mb_slow", 32, 44, 4, 9, ("mb", "slow"), dict(cbca_i2=2, blur_sigma=1.67), 13),
]


def to_np(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    shim = refdriver.ShimLibrary(refdriver.REF_LIB)
    dev = torch.device("cuda:0")
    for name, H, W, C, D, preset, over, seed in CASES:
        opt = pipeline.make_params(*preset, **over)
        pair = synth.make_pair(H, W, C, D, seed=seed)
        x_batch = torch.from_numpy(np.stack([pair["imgL"], pair["imgR"]])[:, None]).to(dev)
        feats = torch.from_numpy(np.stack([pair["featL"], pair["featR"]])).to(dev)
        stages = {}
        refdriver.stereo_predict(shim, x_batch, feats, opt, D, stages=stages)
        torch.cuda.synchronize()
        out = to_np(stages)
        out.update(featL=pair["featL"], featR=pair["featR"], imgL=pair["imgL"], imgR=pair["imgR"])
        out["meta"] = np.array([H, W, C, D], dtype=np.int64)
        out["opt_names"] = np.array([k for k, _ in opt._fields_])
        out["opt_values"] = np.array([float(getattr(opt, k)) for k, _ in opt._fields_], dtype=np.float64)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
        print(name, "ok", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim > 1})

    # stand-alone operators
    rng = np.random.default_rng(21)
    ops = {}
    x = torch.from_numpy(rng.standard_normal((2, 5, 9, 13)).astype(np.float32)).to(dev)
    norm = torch.empty((2, 1, 9, 13), device=dev)
    outn = torch.empty_like(x)
    shim.call("Normalize_forward", x, norm, outn)
    ops.update(norm_in=x, norm_norm=norm, norm_out=outn)
    img0 = torch.from_numpy(rng.standard_normal((1, 1, 14, 22)).astype(np.float32)).to(dev)
    img1 = torch.from_numpy(rng.standard_normal((1, 1, 14, 22)).astype(np.float32)).to(dev)
    for direction in (-1, 1):
        o = torch.empty((1, 7, 14, 22), device=dev)
        shim.call("ad", img0, img1, o, direction)
        ops["ad_%d" % direction] = o
        o2 = torch.empty((1, 7, 14, 22), device=dev)
        shim.call("census", img0, img1, o2, direction)
        ops["census_%d" % direction] = o2
    ops.update(adc_img0=img0, adc_img1=img1)
    vol = torch.from_numpy(rng.standard_normal((2, 6, 5, 7)).astype(np.float32)).to(dev)
    vol[0, 0, 0, 0] = float("nan")
    vol[1, 3, 2, 2] = float("nan")
    am = torch.empty((2, 1, 5, 7), device=dev)
    shim.call("spatial_argmin", vol, am)
    ops.update(argmin_in=vol, argmin_out=am)
    # median2d with several kernel sizes, mean2d with a small kernel
    dimg = torch.from_numpy((rng.random((1, 1, 12, 17)) * 20).astype(np.float32)).to(dev)
    ops["post_img"] = dimg
    for k in (3, 5, 7):
        ops["median_%d" % k] = shim.call("median2d", dimg, k)[0]
    kern = refdriver.gaussian(1.2).to(dev)
    ops["mean2d_kernel"] = kern
    ops["mean2d_out"] = shim.call("mean2d", dimg, kern, 3.0)[0]
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(outdir, "ops.npz"), **to_np(ops))
    print("ops ok")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
