"""Per-opcode and per-instruction warp-stall samples of one kernel from an ncu report
(`ncu --set full --import-source on`), read through `ncu -i ... --page source --csv --print-source sass`.

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [N top instructions]

This is how the class-byte stalls of the SGM scans were found (profiles/r1_sgm_v3_ncu.txt).
"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, body = rows[1], rows[2:]                       # row 0 names the kernel
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]]) for r in body) or 1
    print("kernel:", rows[0][1] if len(rows[0]) > 1 else "?")
    print("total samples", tot, "instructions", len(body))
    samples, long_sb, count, executed = (collections.Counter() for _ in range(4))
    for r in body:
        toks = r[ix["Source"]].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        samples[op] += int(r[ix["# Samples"]])
        long_sb[op] += int(r[ix["stall_long_sb"]])
        count[op] += 1
        executed[op] += int(r[ix["Instructions Executed"]])
    for op, v in samples.most_common(16):
        print("%-12s n=%4d exec %9d samples %6d (%4.1f%%) long_sb %6d" % (op, count[op], executed[op], v, 100.0 * v / tot, long_sb[op]))
    print("--- top instructions")
    for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:ntop]:
        print(r[ix["# Samples"]].rjust(6), "long", r[ix["stall_long_sb"]].rjust(5), "short", r[ix["stall_short_sb"]].rjust(5),
              "wait", r[ix["stall_wait"]].rjust(5), r[ix["Source"]][:90])


if __name__ == "__main__":
    try:
        main()
    except BrokenPipeError:
        pass
