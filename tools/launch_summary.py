"""Summarise an ncu --csv launch list (gpu__time_duration + dram bytes): per kernel name count, mean us, MB read / written.
    python tools/launch_summary.py gpurun_out/x_launches.csv [skip_first_n_launches_per_kernel]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    k = d["Kernel Name"][:70]
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    if u in ("ns", "nsecond"):
        v /= 1e3
    elif u in ("ms", "msecond"):
        v *= 1e3
    elif u in ("s", "second"):
        v *= 1e6
    elif u == "byte":
        v /= 1e6
    elif u == "Kbyte":
        v /= 1e3
    elif u == "Gbyte":
        v *= 1e3
    agg.setdefault(k, {}).setdefault(d["Metric Name"], []).append(v)
tot = 0.0
for k, v in agg.items():
    t = v.get("gpu__time_duration.sum", [0])
    rd = v.get("dram__bytes_read.sum", [0])
    wr = v.get("dram__bytes_write.sum", [0])
    print("%-72s n=%3d  t=%9.1f us  rd=%8.1f MB  wr=%8.1f MB" % (k, len(t), sum(t) / len(t), sum(rd) / len(rd), sum(wr) / len(wr)))

# --json OUT: DRAM bytes per launch (read + write) keyed by bench.py's stage names, for roofline.traffic
if "--json" in sys.argv:
    import json

    def tot(pattern, per_call=1):
        b = 0.0
        for k, v in agg.items():
            if pattern in k:
                rd, wr = v.get("dram__bytes_read.sum", [0]), v.get("dram__bytes_write.sum", [0])
                b += (sum(rd) / len(rd) + sum(wr) / len(wr)) * per_call
        return int(b * 1e6) if b else None

    out = {"StereoJoin": tot("stereo_join"), "cbca_fast": tot("cbca_ws_kernel") or tot("cbca_tma_kernel"), "cbca_exact": tot("cbca_win_kernel"),
           "transpose_in": tot("transpose_rows_kernel"), "transpose_out": tot("transpose_rows_kernel"), "argmin": tot("argmin_pitched_kernel"),
           "sgm2": (tot("sgm_class_kernel") or 0) + (tot("sgm_sel_kernel") or 0) + (tot("sgm_hpair_kernel") or 0) + (tot("sgm_pass_kernel") or 0),
           "source": "ncu launch list %s (dram__bytes_read.sum + dram__bytes_write.sum, mean per launch; sgm2 = class + selector + hpair + the "
                     "two vertical passes of one call)" % sys.argv[1]}
    json.dump({k: v for k, v in out.items() if v}, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
