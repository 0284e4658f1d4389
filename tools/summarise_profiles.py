"""Reads gpurun_out/*.ncu-rep + launches_*.csv (written by tools/gpu_profile.sh on the GPU box) here, without
a GPU, and writes the tracked summaries under profiles/:  python tools/summarise_profiles.py r1"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "l1tex__t_set_conflicts_pipe_lsu_mem_global_op_atom.sum",
    "l1tex__t_set_conflicts_pipe_lsu_mem_global_op_red.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def to_bytes(val: str, unit: str) -> float:
    v = float(val.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


def raw(rep: Path):
    r = list(csv.reader(subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    if len(r) < 3:
        return []
    h, u = r[0], r[1]
    return [{n: (row[h.index(n)], u[h.index(n)]) for n in ["Kernel Name"] + METRICS if n in h} for row in r[2:]]


def main():
    PROF.mkdir(exist_ok=True)
    lf = OUT / f"launches_{tag}.csv"
    if lf.exists():
        lines = [line for line in lf.read_text().splitlines() if not line.startswith("==")]
        per = collections.OrderedDict()
        for row in csv.DictReader(lines):
            if row.get("Metric Name") == "gpu__time_duration.sum":
                name = row["Kernel Name"].split("(")[0]
                v = float(row["Metric Value"].replace(",", ""))
                unit = row["Metric Unit"]
                v = v / 1000 if unit.startswith("n") else (v * 1000 if unit.startswith("m") else v)
                per.setdefault(name, []).append(v)
        tot = sum(sum(v) for v in per.values())
        out = [f"# {tag} launch list: ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120, `python tools/prof_target.py frame 3` (tools/gpu_profile.sh)",
               "# config C2, steady-state frames of the graph-replayed pipeline; per-launch times are cold-cache and serialised: compare SHARES",
               f"{'kernel':45s} {'launches':>8s} {'mean_us':>9s} {'share_%':>8s}"]
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            out.append(f"{k:45s} {len(v):8d} {sum(v) / len(v):9.2f} {100 * sum(v) / tot:8.1f}")
        out.append(f"total {tot:.1f} us over {sum(len(v) for v in per.values())} launches")
        (PROF / f"{tag}_launches_C2.txt").write_text("\n".join(out) + "\n")
        print("\n".join(out))
    traffic = {}
    txt = [f"# {tag}: ncu --set full --clock-control none --import-source on, one launch per report (tools/gpu_profile.sh)"]
    for rep in sorted(OUT.glob(f"prof_*_{tag}.ncu-rep")):
        rows = raw(rep)
        if not rows:
            continue
        row = rows[0]
        txt.append(f"\n## {rep.name}")
        for k, (v, u) in row.items():
            txt.append(f"{k:82s} {v} {u}")
        name = row["Kernel Name"][0].split("(")[0].replace("void ", "").replace("vh::", "")
        if "dram__bytes_read.sum" in row:
            traffic[name] = to_bytes(*row["dram__bytes_read.sum"]) + to_bytes(*row["dram__bytes_write.sum"])
    (PROF / f"{tag}_ncu_summary.txt").write_text("\n".join(txt) + "\n")
    (PROF / f"{tag}_traffic.json").write_text(json.dumps(
        {"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, one ncu --set full capture each (cold L2)", "bytes": traffic}, indent=1) + "\n")
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
