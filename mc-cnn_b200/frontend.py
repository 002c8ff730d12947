"""Image front end of ``main.lua ... -a predict`` (main.lua:1084-1104): load the two views, grey
conversion, standardisation, ``stereo_predict`` on the GPU, and the raw ``.bin`` outputs.

With arch ``ad`` or ``census`` (main.lua:932-942) no trained network is needed, so two PNG files
are enough to run the whole hot path end to end; with arch ``fast`` the caller supplies the
(2,C,H,W) tower output.

Parity notes (SURVEY.md 8c iv): ``image.load`` / ``image.rgb2y`` / ``mean`` / ``std`` are Torch7's
``image`` and ``TH`` (version unpinned, not under /root/reference).  This module restates their
documented arithmetic -- byte values as floats, Y = 0.299 R + 0.587 G + 0.114 B accumulated in
fp32 in that order, mean and the unbiased (n-1) standard deviation accumulated in double, the
scalar applied in fp32 -- **parity unpinned** for the last bit of the double accumulations; every
stage after the (2,1,H,W) batch is the bit-exact path of :mod:`pipeline`.
"""
import os

import numpy as np


def load_image(path):
    """``image.load(path, nil, 'byte'):float()``: (C,H,W) float32 with the byte values 0..255
    (C = 1 for grey files, 3 for colour; an alpha channel is dropped)."""
    from PIL import Image

    with Image.open(path) as im:
        if im.mode in ("I;16", "I;16B", "I"):            # 16-bit grey: 'byte' loading keeps the high byte
            a = (np.asarray(im, dtype=np.uint32) >> 8).astype(np.float32)[None]
        elif im.mode in ("L", "1", "P", "LA"):
            a = np.asarray(im.convert("L"), dtype=np.uint8).astype(np.float32)[None]
        else:
            a = np.asarray(im.convert("RGB"), dtype=np.uint8).astype(np.float32).transpose(2, 0, 1)
    return np.ascontiguousarray(a)


def rgb2y(x):
    """``image.rgb2y``: (3,H,W) -> (1,H,W), Y = 0 + 0.299 R, += 0.587 G, += 0.114 B in fp32."""
    x = np.asarray(x, dtype=np.float32)
    assert x.ndim == 3 and x.shape[0] == 3
    y = np.float32(0.299) * x[0]
    y = y + np.float32(0.587) * x[1]
    y = y + np.float32(0.114) * x[2]
    return y[None].astype(np.float32)


def standardise(x):
    """``x:add(-x:mean()):div(x:std())`` (main.lua:1094-1095): double accumulation, unbiased std taken
    AFTER the mean was subtracted, scalars applied in fp32."""
    x = np.asarray(x, dtype=np.float32)
    n = x.size
    assert n > 1
    m = float(x.astype(np.float64).sum() / n)
    x = x + np.float32(-m)
    m2 = x.astype(np.float64).sum() / n
    var = float(((x.astype(np.float64) - m2) ** 2).sum() / (n - 1))
    assert var > 0, "constant image"
    return (x / np.float32(np.sqrt(var))).astype(np.float32)


def make_batch(left, right):
    """main.lua:1085-1100: two image files (or (C,H,W) / (H,W) arrays) -> x_batch (2,1,H,W) float32."""
    out = []
    for v in (left, right):
        a = load_image(v) if isinstance(v, (str, os.PathLike)) else np.asarray(v, dtype=np.float32)
        if a.ndim == 2:
            a = a[None]
        if a.shape[0] == 3:
            a = rgb2y(a)                                                        # :1088-1092
        assert a.shape[0] == 1, "grey or RGB images expected"
        out.append(standardise(a))                                              # :1094-1095
    assert out[0].shape == out[1].shape, "the two views must have the same size"
    return np.ascontiguousarray(np.stack(out))                                  # :1097-1100


def write_bin(path, t):
    """``torch.DiskFile(path,'w'):binary():writeFloat(t:float():storage())``: raw little-endian fp32"""
    np.ascontiguousarray(t, dtype="<f4").tofile(path)


def predict(left, right, dataset="kitti", arch="census", disp_max=70, features=None, out_dir=None, device=0,
            net_fname=None, **overrides):
    """``main.lua <dataset> <arch> -a predict -net_fname N -left L -right R -disp_max D`` from the image files on:
    returns disp (H,W) float32 (numpy).  With `out_dir`, writes right.bin and left.bin ((1,D,H,W), in
    that order, main.lua:1042-1047) and disp.bin ((1,1,H,W), :1103) there.  arch 'fast' / 'slow' take the trained
    network from `net_fname` (the reference's ascii .t7, main.lua:893-901, read by :mod:`t7`) and run it on the GPU
    (feature tower, and for 'slow' the scorer head); arch 'fast' alternatively accepts `features` (2,C,H,W), the
    unit-norm tower output; 'ad' / 'census' need neither."""
    import torch

    from . import pipeline

    net = None
    if arch in ("fast", "slow") and features is None:                          # decided before any GPU work
        if net_fname is None:
            raise ValueError("arch '%s' needs -net_fname (or, for 'fast', the tower output in `features`)" % arch)
        from . import t7

        net = t7.load_net(net_fname)                                            # main.lua:893-901
        if net.arch != arch:
            raise ValueError("%s holds a '%s' network, arch '%s' was asked for" % (net_fname, net.arch, arch))
    elif arch == "slow":
        raise ValueError("arch 'slow' needs -net_fname: the scorer head's weights come with the tower's")
    elif arch not in ("fast", "ad", "census"):
        raise ValueError("arch must be 'fast', 'slow', 'ad' or 'census'")

    batch = make_batch(left, right)
    dev = torch.device("cuda", device)
    x_batch = torch.from_numpy(batch).to(dev)
    feats = head = None
    if net is not None:
        from .feature_tower import FeatureTower
        from .scorer_head import ScorerHead

        with torch.cuda.device(dev):
            feats = FeatureTower(net.tower, arch=arch, device=dev.index).forward(x_batch)      # main.lua:944 / :957
            if arch == "slow":
                head = ScorerHead(net.head, device=dev.index)                                  # net_te2, main.lua:896
    elif arch == "fast":
        feats = torch.as_tensor(features, dtype=torch.float32).to(dev).contiguous()
    opt = pipeline.make_params(dataset, arch, **overrides)
    with torch.cuda.device(dev):
        disp, volL, volR = pipeline.stereo_predict(x_batch, feats, opt, int(disp_max), want_vols=True, arch=arch, head=head)
    disp_h = disp.cpu().numpy()
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        write_bin(os.path.join(out_dir, "right.bin"), volR.cpu().numpy())       # direction 1 comes first
        write_bin(os.path.join(out_dir, "left.bin"), volL.cpu().numpy())
        write_bin(os.path.join(out_dir, "disp.bin"), disp_h)
    return disp_h[0, 0]
