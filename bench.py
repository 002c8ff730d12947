#!/usr/bin/env python
"""bench.py -- stereo pairs/s of the hot path (StereoJoin -> CBCA -> SGM -> post) on N B200s.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`, one JSON
line on stdout from rank 0.  A "step" = one stereo pair (BASELINE.json config 3: 370x1226, d=228,
C=64 features, kitti-slow post-processing with CBCA x4 + 4-direction SGM) through the fused native
pipeline, one pair per GPU per step (pairs shard over GPUs with no collective: weak scaling).

  value      whole-job pairs/s with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same through the host-buffer C-ABI call (H2D of features+images and D2H of disp inside)
  roofline   the dominant kernel of the step, timed alone with CUDA events on its stream
  cpu_baseline  the CPU oracle (oracle/, a port of the reference) on a bounded sample, rank 0, N=1

`--impl reference` times the reference's OWN implementation of the path: adcensus.cu compiled
unmodified (oracle/_ref/libadcensus_ref.so) and driven in main.lua's order.  The reference has no
CPU path (every op takes torch.CudaTensor), so this arm also runs on the B200 (BASELINE.json
north_star says so); rank 0 only.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json config 3: the configuration the metric is quoted on (default)
    "k228": dict(name="kitti_accurate_370x1226_d228_cbca4_sgm4", H=370, W=1226, D=228, C=64,
                 preset=("kitti", "accurate_cbca4"),
                 desc="kitti slow post-processing, cbca_i1=2 cbca_i2=2, sgm2 (4 directions), LR check, subpixel, "
                      "median5, bilateral"),
    # BASELINE.json config 2 (reported in profiles/, not the driver's bench line)
    "k70": dict(name="kitti_fast_370x1226_d70", H=370, W=1226, D=70, C=64, preset=("kitti", "fast"),
                desc="kitti fast preset (no CBCA), sgm2 (4 directions), LR check, subpixel, median5, bilateral"),
}
WORKLOAD = WORKLOADS["k228"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--stages", action="store_true", help="also print a per-stage timing table to stderr")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--small", action="store_true", help="tiny workload (debug only; not a valid bench line)")
    ap.add_argument("--workload", default="k228", choices=sorted(WORKLOADS))
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier_sync(world):
    import torch
    import torch.distributed as dist

    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch
    import torch.distributed as dist

    if world == 1:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_inputs(cfg, n_pairs, seed0):
    import numpy as np
    import torch

    from mccnn_b200 import synth

    pairs = []
    for i in range(n_pairs):
        p = synth.make_pair(cfg["H"], cfg["W"], cfg["C"], cfg["D"], seed=seed0 + i)
        pairs.append({k: torch.from_numpy(np.ascontiguousarray(p[k])).pin_memory() for k in ("featL", "featR", "imgL", "imgR")})
    return pairs


def algorithmic_bytes(cfg):
    H, W, D, C = cfg["H"], cfg["W"], cfg["D"], cfg["C"]
    V = 4 * D * H * W
    F = 4 * C * H * W
    valid = 4 * H * (D * W - D * (D - 1) // 2)
    return {
        "StereoJoin": 2 * F + 2 * valid,
        "cbca": 2 * V + 32 * H * W,               # per iteration
        "sgm2": 11 * V,                           # documented 4-pass design: 4 x (read in + RMW out) - 1 read
        "transpose": 2 * V,
        "argmin": V + 4 * H * W,
        "fill_nan": 2 * V,
    }


def time_op(fn, iters, flush):
    """average CUDA-event duration (ms) of fn() on torch's current stream, L2 flushed before each"""
    import torch

    fn()            # untimed: lazy kernel load, pool growth
    evs = []
    for _ in range(iters):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / len(evs)


def stage_table(cfg, opt, dev_in, iters=5):
    """per-operator timings at the bench workload through the op-level C ABI (isolated, L2 flushed)"""
    import torch

    from mccnn_b200 import adcensus

    H, W, D, C = cfg["H"], cfg["W"], cfg["D"], cfg["C"]
    dev = dev_in["featL"].device
    flushbuf = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    flush = lambda: flushbuf.fill_(1.0)
    fL, fR = dev_in["featL"][None], dev_in["featR"][None]
    iL, iR = dev_in["imgL"][None], dev_in["imgR"][None]
    vols = torch.empty((2, D, H, W), device=dev)
    adcensus.fill_nan(vols)
    t = {}
    t["fill_nan"] = time_op(lambda: adcensus.fill_nan(vols), iters, flush)
    t["StereoJoin"] = time_op(lambda: adcensus.StereoJoin(fL, fR, vols[0:1], vols[1:2]), iters, flush)
    x0c = torch.empty((1, 4, H, W), device=dev); x1c = torch.empty((1, 4, H, W), device=dev)
    t["cross"] = time_op(lambda: adcensus.cross(iL, x0c, opt.L1, opt.tau1), iters, flush)
    adcensus.cross(iR, x1c, opt.L1, opt.tau1)
    tmp = torch.empty((1, D, H, W), device=dev)
    # arms packed once (as the pipeline does), so that the timed launch is the aggregation kernel alone
    lib = adcensus.lib()
    lib.mccnn_packed_arms_bytes.restype = ctypes.c_size_t
    packed = torch.empty(lib.mccnn_packed_arms_bytes(H, W), dtype=torch.uint8, device=dev)
    vp = lambda t_: ctypes.c_void_p(t_.data_ptr())
    assert lib.mccnn_pack_arms(vp(x0c), vp(x1c), vp(packed), H, W, adcensus._stream(x0c)) == 0
    t["cbca"] = time_op(lambda: lib.mccnn_cbca_packed(vp(packed), vp(x0c), vp(x1c), vp(vols[0:1]), vp(tmp), D, H, W, -1,
                                                       max(opt.L1, 2), adcensus._stream(tmp)), iters, flush)
    volt = adcensus.transpose_dhw_to_hwd(vols[0:1])
    t["transpose"] = time_op(lambda: adcensus.lib().mccnn_transpose_dhw_to_hwd(
        adcensus._t(vols[0:1], 1, "t"), adcensus._t(volt, 2, "t"), D, H, W, adcensus._stream(volt)), iters, flush)
    out = torch.zeros_like(volt)
    t["sgm2"] = time_op(lambda: adcensus.sgm2(iL, iR, volt, out, None, opt.pi1, opt.pi2, opt.tau_so, opt.alpha1,
                                             opt.sgm_q1, opt.sgm_q2, -1), iters, flush)
    t["argmin"] = time_op(lambda: adcensus.argmin(tmp), iters, flush)
    d = adcensus.argmin(tmp)
    d1 = adcensus.argmin(vols[1:2])
    outl = torch.zeros_like(d)
    t["outlier_detection"] = time_op(lambda: adcensus.outlier_detection(d, d1, outl, D), iters, flush)
    t["interpolate_occlusion"] = time_op(lambda: adcensus.interpolate_occlusion(d, outl), iters, flush)
    t["interpolate_mismatch"] = time_op(lambda: adcensus.interpolate_mismatch(d, outl), iters, flush)
    t["subpixel"] = time_op(lambda: adcensus.subpixel_enchancement(d, tmp, D), iters, flush)
    t["median2d"] = time_op(lambda: adcensus.median2d(d, 5), iters, flush)
    kern = adcensus.gaussian(opt.blur_sigma).to(dev)
    t["mean2d"] = time_op(lambda: adcensus.mean2d(d, kern, opt.blur_t), iters, flush)
    return t


def cpu_baseline(cfg, opt):
    """the CPU oracle (a port: the reference ships no CPU path) on a bounded row band"""
    import numpy as np

    from mccnn_b200 import synth
    from oracle import oracle as orc

    rows = min(cfg["H"], 96)
    p = synth.make_pair(rows, cfg["W"], cfg["C"], cfg["D"], seed=5)
    op = orc.Params(**opt.as_dict())
    orc.lib()
    t0 = time.perf_counter()
    orc.stereo_predict(p["featL"], p["featR"], p["imgL"], p["imgR"], cfg["D"], op)
    dt = time.perf_counter() - t0
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    frac = rows / cfg["H"]
    return {"value": frac / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": "rows 0..%d of one %dx%d d=%d pair (%.1f%% of a pair) in %.1f s, OpenMP oracle; scaled by rows" % (
                rows - 1, cfg["H"], cfg["W"], cfg["D"], 100 * frac, dt)}


def run_b200(args):
    import torch

    import mccnn_b200  # noqa: F401
    from mccnn_b200 import adcensus, pipeline

    # keep stdout clean for the one JSON line (NCCL / torch may print banners there)
    json_fd = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = dist_setup(args)
    cfg = dict(WORKLOADS[args.workload])
    if args.small:
        cfg.update(name="debug_small", H=64, W=128, D=16)
    opt = pipeline.make_params(*cfg["preset"])
    dev = torch.device("cuda", local)
    adcensus.lib()  # fail loudly if the CUDA library is missing

    pairs = make_inputs(cfg, 2, 1000 + 16 * rank)
    dev_in = [{k: v.to(dev) for k, v in p.items()} for p in pairs]
    sp = pipeline.StereoPipeline(cfg["C"], cfg["D"], cfg["H"], cfg["W"], opt, device=local)
    disp = torch.empty((cfg["H"], cfg["W"]), device=dev)
    K, Wm = args.steps, max(args.warmup, 3)

    # ---- device-resident throughput -------------------------------------------------------
    for i in range(Wm):
        x = dev_in[i % 2]
        sp.run(x["featL"], x["featR"], x["imgL"], x["imgR"], disp=disp)
    barrier_sync(world)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        x = dev_in[i % 2]
        sp.run(x["featL"], x["featR"], x["imgL"], x["imgR"], disp=disp)
    e1.record()
    barrier_sync(world)
    ms_total = max_over_ranks(e0.elapsed_time(e1), world)
    clocks = sampler.stop()
    launches = sp.launches_per_run * K

    # ---- same, with the opt-in approximate CBCA (NOT bit-exact; reported beside the headline) ----
    sp.set_fast_cbca(True)
    for i in range(2):
        x = dev_in[i % 2]
        sp.run(x["featL"], x["featR"], x["imgL"], x["imgR"], disp=disp)
    barrier_sync(world)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(K):
        x = dev_in[i % 2]
        sp.run(x["featL"], x["featR"], x["imgL"], x["imgR"], disp=disp)
    f1.record()
    barrier_sync(world)
    ms_fast = max_over_ranks(f0.elapsed_time(f1), world)
    disp_fast = disp.clone()
    sp.set_fast_cbca(False)
    sp.run(dev_in[(K - 1) % 2]["featL"], dev_in[(K - 1) % 2]["featR"], dev_in[(K - 1) % 2]["imgL"], dev_in[(K - 1) % 2]["imgR"], disp=disp)
    torch.cuda.synchronize()
    fast_diff_frac = float(((disp_fast - disp).abs() > 1e-4 * disp.abs().clamp(min=1.0)).float().mean().item())

    # ---- end to end through the host-buffer C-ABI call --------------------------------------
    # one batch call per timed region: every step's inputs cross PCIe from pinned host memory and every
    # step's disparity map comes back; copies of neighbouring steps overlap the kernels (3 streams)
    host_pairs = [tuple(pairs[i % 2][k] for k in ("featL", "featR", "imgL", "imgR")) for i in range(K)]
    disps_h = [torch.empty((cfg["H"], cfg["W"]), dtype=torch.float32).pin_memory() for _ in range(K)]
    sp.run_host_batch(host_pairs[:2], disps_h[:2])
    barrier_sync(world)
    t0 = time.perf_counter()
    sp.run_host_batch(host_pairs, disps_h)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    barrier_sync(world)
    # single-pair latency through the same boundary (no overlap possible)
    t1 = time.perf_counter()
    sp.run_host(*host_pairs[0], disp=disps_h[0])
    e2e_single_ms = (time.perf_counter() - t1) * 1e3
    F = 4 * cfg["C"] * cfg["H"] * cfg["W"]
    I = 4 * cfg["H"] * cfg["W"]

    # ---- per-stage timings + roofline of the dominant kernel (rank 0) -------------------------
    out = None
    if rank == 0:
        stages = stage_table(cfg, opt, dev_in[0], iters=max(3, min(K, 10)))
        if args.stages:
            for k, v in stages.items():
                sys.stderr.write("%-24s %8.3f ms\n" % (k, v))
        ab = algorithmic_bytes(cfg)
        n_cbca = 2 * (opt.cbca_i1 + opt.cbca_i2)
        share = {"StereoJoin": stages["StereoJoin"], "cbca": n_cbca * stages["cbca"], "sgm2": 2 * opt.sgm_i * stages["sgm2"],
                 "transpose": 4 * stages["transpose"], "argmin": 2 * stages["argmin"], "fill_nan": stages["fill_nan"]}
        dom = max(share, key=share.get)
        peak, peak_src = load_peaks()

        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "r1_traffic_k228.json")
        if args.workload == "k228" and not args.small and os.path.exists(tpath):
            traffic = json.load(open(tpath))

        def roof(name):
            ach = ab[name] / (stages[name] * 1e-3) / 1e9
            return {"kernel": name, "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": traffic.get(name), "peak_source": peak_src,
                    "algorithmic_bytes": ab[name], "ms": round(stages[name], 4)}

        out = {
            "metric": "stereo pairs/sec (370x1226 d=%d)" % cfg["D"], "value": round(world * K / (ms_total * 1e-3), 3),
            "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": round(ms_total / K, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "H": cfg["H"], "W": cfg["W"], "D": cfg["D"], "C": cfg["C"],
                       "preset": cfg.get("desc", ""),
                       "pairs_per_gpu_per_step": 1, "parallelism": "pairs sharded over GPUs, no collective",
                       "schedule": "the two directions of a pair overlapped on two streams (mccnn_pipeline_set_overlap mode 2)",
                       "l2": "per-step working set (0.23 GB features + 1.65 GB volumes) exceeds the 126 MB L2; "
                             "stage timings flush L2 with a 256 MB write"},
            "clocks": clocks,
            "e2e": {"value": round(world * K / e2e_s, 3), "unit": "pairs/s", "h2d_bytes_per_step": 2 * F + 2 * I,
                    "d2h_bytes_per_step": I, "api": "mccnn_pipeline_run_host_batch (host buffers in, host disparity maps out)",
                    "single_pair_latency_ms": round(e2e_single_ms, 3)},
            "gpu_launches": launches,
            "fast_cbca": {"value": round(world * K / (ms_fast * 1e-3), 3), "unit": "pairs/s", "ms_per_step": round(ms_fast / K, 4),
                          "note": "opt-in prefix-sum CBCA (mccnn_pipeline_set_fast_cbca): volumes within ~1e-6 relative of the "
                                  "exact mode, not bit-exact; the headline `value` uses the exact mode",
                          "disp_pixels_differing_from_exact_frac": fast_diff_frac},
            "roofline": roof(dom),
            "roofline_stereojoin": roof("StereoJoin"),
            "roofline_cbca": roof("cbca"),
            "stage_ms": {k: round(v, 4) for k, v in stages.items()},
            "step_share_ms": {k: round(v, 4) for k, v in share.items()},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(cfg, opt)
    sp.close()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        os.write(json_fd, (json.dumps(out) + "\n").encode())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the reference arm
    from oracle import refdriver

    if not os.path.exists(refdriver.REF_LIB):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libadcensus_ref.so missing (reference tree was not present at build time)"}))
        return
    import numpy as np
    import torch

    import mccnn_b200  # noqa: F401
    from mccnn_b200 import pipeline

    torch.cuda.set_device(0)
    cfg = dict(WORKLOADS[args.workload])
    if args.small:
        cfg.update(name="debug_small", H=64, W=128, D=16)
    opt = pipeline.make_params(*cfg["preset"])
    dev = torch.device("cuda", 0)
    pairs = make_inputs(cfg, 2, 1000)
    shim = refdriver.ShimLibrary(refdriver.REF_LIB)
    xb = [torch.stack([p["imgL"], p["imgR"]])[:, None].to(dev) for p in pairs]
    ft = [torch.stack([p["featL"], p["featR"]]).to(dev) for p in pairs]
    K, Wm = args.steps, max(args.warmup, 1)
    for i in range(Wm):
        refdriver.stereo_predict(shim, xb[i % 2], ft[i % 2], opt, cfg["D"])
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        refdriver.stereo_predict(shim, xb[i % 2], ft[i % 2], opt, cfg["D"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    v = round(K / (ms * 1e-3), 4)
    print(json.dumps({
        "impl": "reference", "metric": "stereo pairs/sec (370x1226 d=%d)" % cfg["D"], "value": v, "unit": "pairs/s",
        "n_gpus": 1, "steps": K, "warmup": Wm, "ms_per_step": round(ms / K, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "H": cfg["H"], "W": cfg["W"], "D": cfg["D"], "C": cfg["C"],
                   "note": "reference adcensus.cu compiled unmodified for sm_100a (oracle/_ref), driven in main.lua:929-1082 "
                           "order with torch ops for the cutorch-side fill/copy/transpose/div; the reference has no CPU path, "
                           "so its own implementation of the path is this GPU one (1 B200, rank 0)"},
        "clocks": clocks,
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": 0, "kind": "reference",
                         "sample": "%d full pairs on one B200 (CUDA reference; host cores only launch)" % K},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
