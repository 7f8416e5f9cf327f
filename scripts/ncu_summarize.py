"""Summarise .ncu-rep captures (ncu --set full) into profiles/: one JSON with per-kernel averages + a markdown table.
usage: python scripts/ncu_summarize.py <tag> <rep> [<rep> ...]   (run where `ncu` is installed; no GPU needed)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_ncu_peak",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__waves_per_multiprocessor": "waves_per_sm",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_throttle",
}
SCALE = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].strip()
        if name.startswith("void "):
            name = name[5:]
        d = {"kernel": name.replace("(bool)", "")}   # e.g. ac_adj_kernel<1>, el_vel_adj<1>
        for k, name in KEYS.items():
            if k in idx and r[idx[k]] != "":
                v = float(r[idx[k]].replace(",", ""))
                v *= SCALE.get(units[idx[k]], 1.0) if ("bytes" in name or name == "duration_us") else 1.0
                d[name] = v
        res.append(d)
    return res


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    launches = [d for r in reps for d in load(r)]
    kernels = {}
    for d in launches:
        kernels.setdefault(d["kernel"], []).append(d)
    summ = {}
    for k, ds in kernels.items():
        avg = {key: sum(x.get(key, 0.0) for x in ds) / len(ds) for key in ds[0] if key != "kernel"}
        avg["launches_captured"] = len(ds)
        avg["dram_bytes_per_launch"] = avg.get("dram_read_bytes", 0) + avg.get("dram_write_bytes", 0)
        avg["dram_GBps"] = avg["dram_bytes_per_launch"] / (avg["duration_us"] * 1e-6) / 1e9
        summ[k] = avg
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    out = dict(tag=tag, source=[os.path.basename(r) for r in reps],
               how="ncu --set full --clock-control none --import-source on (one GPU, under gpurun); averages over the "
                   "captured launches; durations are serialised/cold-cache, use them for shares only", kernels=summ)
    json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_summary_%s.json" % tag), "w"), indent=1)
    json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_summary.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", "ncu_summary_%s.md" % tag), "w") as f:
        f.write("| kernel | us | DRAM MB/launch | DRAM GB/s | regs | occupancy % | fp64 pipe % | L2 hit % | long-scoreboard stall |\n|---|---|---|---|---|---|---|---|---|\n")
        for k, a in summ.items():
            f.write("| %s | %.1f | %.1f | %.0f | %d | %.1f | %.1f | %.1f | %.2f |\n" % (
                k, a["duration_us"], a["dram_bytes_per_launch"] / 1e6, a["dram_GBps"], a.get("registers", 0),
                a.get("achieved_occupancy_pct", 0), a.get("fp64_pipe_pct", 0), a.get("l2_hit_pct", 0),
                a.get("stall_long_scoreboard", 0)))
    print(open(os.path.join(ROOT, "profiles", "ncu_summary_%s.md" % tag)).read())


if __name__ == "__main__":
    main()
