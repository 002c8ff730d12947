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
