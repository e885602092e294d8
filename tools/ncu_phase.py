"""Sums ncu per-line instruction / sample shares over source line ranges.
Usage: python tools/ncu_phase.py report.ncu-rep file.cu 'name:lo-hi' ..."""
import csv
import subprocess
import sys


def main():
    rep, src = sys.argv[1], sys.argv[2]
    ranges = []
    for a in sys.argv[3:]:
        nm, r = a.split(":")
        lo, hi = r.split("-")
        ranges.append((nm, int(lo), int(hi)))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi_ = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi_]
    si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    num = lambda x: int(x) if x.lstrip("-").isdigit() else 0
    agg = {}
    for r in rows[hi_ + 1:]:
        if len(r) > ti and r[0].isdigit():
            a = agg.setdefault(int(r[0]), [0, 0, 0])
            a[0] += num(r[si]); a[1] += num(r[ii]); a[2] += num(r[ti])
    ts = sum(a[0] for a in agg.values()) or 1
    tin = sum(a[1] for a in agg.values()) or 1
    print("total samples %d warp-inst %d" % (ts, tin))
    for nm, lo, hi in ranges:
        s = sum(a[0] for l, a in agg.items() if lo <= l <= hi)
        i = sum(a[1] for l, a in agg.items() if lo <= l <= hi)
        t = sum(a[2] for l, a in agg.items() if lo <= l <= hi)
        print("%-14s lines %4d-%4d  samples %5.1f%%  inst %5.1f%% (%.3e)  lanes %4.1f" % (nm, lo, hi, 100 * s / ts, 100 * i / tin, i, t / max(i, 1)))


if __name__ == "__main__":
    main()
