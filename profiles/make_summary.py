"""Turns `ncu --set full` reports (gpurun_out/*.ncu-rep) into the small JSON summaries committed in this directory.
Usage: python profiles/make_summary.py gpurun_out/k1.ncu-rep gpurun_out/k2.ncu-rep > profiles/rNN_ncu_summary.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active"]


def summarize(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        e = {"kernel": d.get("Kernel Name", "?")}
        for k in KEYS:
            if k in d:
                e[k] = f"{d[k]} {u[k]}".strip()
        stalls = {h: float(d[h]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h]}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        e["top_stalls_per_issue"] = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(v, 3) for k, v in top}
        out.append(e)
    return out


if __name__ == "__main__":
    res = []
    for p in sys.argv[1:]:
        res += summarize(p)
    print(json.dumps(res, indent=1))
