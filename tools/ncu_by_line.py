"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.
Usage: python tools/ncu_by_line.py report.ncu-rep [source.cu] [top]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    srcfile = sys.argv[2] if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi]
    si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")

    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0

    agg = {}
    cur_file = None
    for r in rows[hi + 1:]:
        if len(r) <= ti:
            if r and r[0] == "File Path" or (r and "File" in r[0]):
                cur_file = r[1] if len(r) > 1 else None
            continue
        if r[0] and r[0].isdigit():
            a = agg.setdefault((cur_file, int(r[0]), r[1]), [0, 0, 0])
            a[0] += num(r[si]); a[1] += num(r[ii]); a[2] += num(r[ti])
    tots = sum(a[0] for a in agg.values()) or 1
    toti = sum(a[1] for a in agg.values()) or 1
    print("samples", tots, "warp-instructions", toti)
    for (f, ln, txt), (s, i, t) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print("%4d samp %5.1f%% inst %5.1f%% lanes %4.1f | %s" % (ln, 100 * s / tots, 100 * i / toti, t / max(i, 1), txt.strip()[:90]))


if __name__ == "__main__":
    main()
