"""Drive a shim-built Lua operator library on the GPU -- TEST INFRASTRUCTURE.

``ShimLibrary`` loads an .so that was linked with oracle/refshim (the reference's
own adcensus.cu -> oracle/_ref/libadcensus_ref.so, or the shim build of OUR Lua
face -> oracle/_ref/libadcensus_luaface.so), runs ``luaopen_libadcensus`` and lets
Python call ``adcensus.<name>(...)`` with torch CUDA tensors and numbers, exactly
as main.lua does.  ``stereo_predict`` restates main.lua:929-1082 on top of it,
with torch ops standing in for the cutorch ops of the Lua side (fill, copy,
transpose, div, min).

Used by: tests (parity of the CUDA path against the reference itself),
oracle/make_golden.py (fixtures), bench.py --impl reference.  Never by the product.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libadcensus_ref.so")
LUAFACE_LIB = os.path.join(_HERE, "_ref", "libadcensus_luaface.so")


class ShimError(RuntimeError):
    pass


class ShimLibrary:
    def __init__(self, path=REF_LIB):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        l = self.lib
        l.shim_state_new.restype = ctypes.c_void_p
        l.shim_last_error.restype = ctypes.c_char_p
        l.shim_function_name.restype = ctypes.c_char_p
        l.shim_result_tensor.restype = ctypes.c_void_p
        l.shim_result_number.restype = ctypes.c_double
        self.L = ctypes.c_void_p(l.shim_state_new())
        rc = l.luaopen_libadcensus(self.L)  # adcensus.cu:2100-2105
        if rc != 1:
            raise ShimError("luaopen_libadcensus returned %d" % rc)
        l.shim_reset(self.L)

    def functions(self, table="adcensus"):
        n = self.lib.shim_num_functions(table.encode())
        return [self.lib.shim_function_name(table.encode(), i).decode() for i in range(n)]

    def call(self, name, *args, table="adcensus"):
        """adcensus.<name>(*args).  Tensors: CUDA float32 contiguous (or CPU float32 for the
        torch.FloatTensor ops).  Returns the list of Lua results (new CUDA tensors / numbers)."""
        l, L = self.lib, self.L
        l.shim_reset(L)
        keep = []
        for a in args:
            if isinstance(a, torch.Tensor):
                assert a.dtype == torch.float32 and a.is_contiguous()
                sizes = (ctypes.c_long * a.dim())(*a.shape)
                keep.append(sizes)
                if a.is_cuda:
                    l.shim_push_cuda_tensor(L, ctypes.c_void_p(a.data_ptr()), a.dim(), sizes)
                else:
                    l.shim_push_float_tensor(L, ctypes.c_void_p(a.data_ptr()), a.dim(), sizes)
            elif isinstance(a, str):
                l.shim_push_string(L, a.encode())
            else:
                l.shim_push_number(L, ctypes.c_double(float(a)))
        nret = l.shim_call(L, table.encode(), name.encode())
        if nret < 0:
            msg = l.shim_last_error(L).decode()
            l.shim_reset(L)
            raise ShimError(msg)
        results = []
        for i in range(nret):
            if l.shim_result_is_tensor(L, i):
                nd = ctypes.c_int()
                sizes = (ctypes.c_long * 8)()
                l.shim_result_tensor(L, i, ctypes.byref(nd), sizes)
                shape = tuple(sizes[k] for k in range(nd.value))
                t = torch.empty(shape, device="cuda", dtype=torch.float32)
                rc = l.shim_copy_result_to(L, i, ctypes.c_void_p(t.data_ptr()), ctypes.c_size_t(t.numel() * 4))
                if rc != 0:
                    raise ShimError("copy of result %d failed (%d)" % (i, rc))
                results.append(t)
            else:
                results.append(l.shim_result_number(L, i))
        l.shim_reset(L)
        return results


def gaussian(sigma):
    """main.lua:528-540 (double math, stored as float by :cuda())"""
    import math

    kr = int(math.ceil(sigma * 3))
    ks = kr * 2 + 1
    k = torch.empty((ks, ks), dtype=torch.float64)
    for i in range(ks):
        for j in range(ks):
            y, x = i - kr, j - kr
            k[i, j] = math.exp(-(x * x + y * y) / (2 * sigma * sigma))
    return k.float()


def fix_border(vol, n, direction):
    """main.lua:922-927 on a (1,D,H,W) tensor (Torch negative indices count from the end)"""
    W = vol.size(3)
    for i in range(1, n + 1):
        if direction > 0:
            vol[:, :, :, i - 1].copy_(vol[:, :, :, n])
        else:
            vol[:, :, :, W - i].copy_(vol[:, :, :, W - n - 1])


def stereo_predict(shim, x_batch, features, opt, disp_max, want_vols=False, stages=None, directions=(1, -1)):
    """main.lua:929-1082 (arch 'fast') with every adcensus.* call going to `shim`.

    `stages`, when a dict, receives clones of the intermediate tensors (for fixtures).  `directions`: main.lua:954-955 runs
    {1, -1}, or {-1} alone for dataset 'mb' outside `-a predict` (then the right volume is never aggregated).
    """
    call = shim.call
    H, W = x_batch.size(2), x_batch.size(3)
    dev = x_batch.device
    rec = (lambda k, v: stages.__setitem__(k, v.clone())) if stages is not None else (lambda k, v: None)

    vols = torch.full((2, disp_max, H, W), float("nan"), device=dev, dtype=torch.float32)   # :946
    call("StereoJoin", features[0:1], features[1:2], vols[0:1], vols[1:2])                  # :947
    rec("sj_left", vols[0:1]); rec("sj_right", vols[1:2])
    fix_border(vols[0:1], opt.border, -1)                                                   # :948
    fix_border(vols[1:2], opt.border, 1)                                                    # :949
    disp = {}
    out_vols = {}
    vol = None
    for direction in directions:                                                            # :954-955
        tag = "L" if direction == -1 else "R"
        vol = vols[0:1] if direction == -1 else vols[1:2]                                   # :986
        x0c = torch.empty((1, 4, H, W), device=dev, dtype=torch.float32)
        x1c = torch.empty((1, 4, H, W), device=dev, dtype=torch.float32)
        call("cross", x_batch[0], x0c, opt.L1, opt.tau1)                                    # :995
        call("cross", x_batch[1], x1c, opt.L1, opt.tau1)                                    # :996
        rec("x0c", x0c); rec("x1c", x1c)
        tmp_cbca = torch.empty_like(vol)
        for _ in range(opt.cbca_i1):                                                        # :998-1001
            call("cbca", x0c, x1c, vol, tmp_cbca, direction)
            vol.copy_(tmp_cbca)
        rec("cbca1_" + tag, vol)
        volt = vol.transpose(1, 2).transpose(2, 3).clone(memory_format=torch.contiguous_format)                                  # :1008
        out = torch.empty_like(volt)
        tmp = torch.empty((volt.size(2), volt.size(3)), device=dev, dtype=torch.float32)    # :1012
        for _ in range(opt.sgm_i):
            out.zero_()                                                                     # :1014
            call("sgm2", x_batch[0], x_batch[1], volt, out, tmp, opt.pi1, opt.pi2, opt.tau_so,
                 opt.alpha1, opt.sgm_q1, opt.sgm_q2, direction)                             # :1015
            volt.copy_(out).div_(4)                                                         # :1017
        vol.copy_(out.transpose(2, 3).transpose(1, 2)).div_(4)                              # :1019-1020
        rec("sgm_" + tag, vol)
        for _ in range(opt.cbca_i2):                                                        # :1035-1038
            call("cbca", x0c, x1c, vol, tmp_cbca, direction)
            vol.copy_(tmp_cbca)
        if want_vols or stages is not None:
            out_vols[direction] = vol.clone()
        # torch.min(vol, 2) is cutorch; the in-repo statement of its semantics is spatial_argmin
        d = torch.empty((1, 1, H, W), device=dev, dtype=torch.float32)
        call("spatial_argmin", vol, d)                                                      # :1049
        disp[1 if direction == 1 else 2] = d.add_(-1)                                       # :1050
    if 1 in disp:
        rec("disp_R", disp[1])
    rec("disp_L", disp[2])
    d = disp[2]
    if opt.lr_check:                                                                        # :1054-1066
        outlier = torch.zeros_like(d)
        call("outlier_detection", disp[2], disp[1], outlier, disp_max)
        rec("outlier", outlier)
        d = call("interpolate_occlusion", d, outlier)[0]
        rec("occ", d)
        d = call("interpolate_mismatch", d, outlier)[0]
        rec("mis", d)
    d = call("subpixel_enchancement", d, vol, disp_max)[0]                                  # :1068
    rec("subpixel", d)
    d = call("median2d", d, 5)[0]                                                           # :1073
    rec("median", d)
    d = call("mean2d", d, gaussian(opt.blur_sigma).to(dev), opt.blur_t)[0]                  # :1078
    rec("disp", d)
    if want_vols:
        return d, out_vols[-1], out_vols.get(1)
    return d
