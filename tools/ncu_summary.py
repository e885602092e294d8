"""One row of key metrics per `ncu --set full` report (raw page), as markdown + JSON.
Usage: python tools/ncu_summary.py out_prefix report1.ncu-rep [report2 ...]"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "time",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_inst",
    "smsp__inst_executed.sum": "warp_inst",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_sb",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_sb",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_throttle",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
        "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"}
    for i, k in enumerate(hdr):
        if k in KEYS:
            try:
                v = float(vals[i].replace(",", ""))
            except ValueError:
                continue
            d[KEYS[k]] = v * UNIT.get(units[i], 1.0)
    if "time" in d and d["time"] > 0:
        d["dram_GBps"] = (d.get("dram_read", 0) + d.get("dram_write", 0)) / d["time"] / 1e9
    return d


def main():
    prefix = sys.argv[1]
    res = {}
    for rep in sys.argv[2:]:
        try:
            res[rep.split("/")[-1]] = load(rep)
        except Exception as e:  # noqa: BLE001
            res[rep.split("/")[-1]] = {"error": str(e)}
    json.dump(res, open(prefix + ".json", "w"), indent=1)
    cols = ["kernel", "time", "dram_read", "dram_write", "dram_GBps", "dram_pct", "issue_active_pct", "warps_active_pct",
            "lanes_per_inst", "pipe_alu_pct", "pipe_fma_pct", "pipe_lsu_pct", "l2_hit_pct", "regs"]
    with open(prefix + ".md", "w") as f:
        f.write("| report | " + " | ".join(cols) + " |\n|" + "---|" * (len(cols) + 1) + "\n")
        for rep, d in res.items():
            def fmt(c):
                v = d.get(c)
                if v is None:
                    return ""
                if c == "time":
                    return "%.3f ms" % (v * 1e3)
                if c in ("dram_read", "dram_write"):
                    return "%.1f MB" % (v / 1e6)
                if isinstance(v, float):
                    return "%.1f" % v
                return str(v)[:40]
            f.write("| " + rep + " | " + " | ".join(fmt(c) for c in cols) + " |\n")
    print(open(prefix + ".md").read())


if __name__ == "__main__":
    main()
