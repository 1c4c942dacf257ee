#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python profiles/summarize.py <tag>      # e.g. r01

Reads gpurun_out/prof_*_<tag>.ncu-rep (ncu --set full) and gpurun_out/launches_*_<tag>.csv
(ncu --metrics gpu__time_duration.sum) and writes profiles/<tag>_*.txt.
"""
import collections
import csv
import io
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def full_report(path, dest):
    raw = subprocess.run(["ncu", "-i", str(path), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# {path.name}: ncu --set full --clock-control none (per launch; cold-ish caches, serialised)"]
    for r in rows[2:]:
        lines.append("")
        lines.append(f"kernel: {r[hdr.index('Kernel Name')]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                lines.append(f"  {m:72s} {r[i]} {units[i]}")
    dest.write_text("\n".join(lines) + "\n")
    print("wrote", dest)


def launch_list(path, dest):
    text = path.read_text()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    per = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = val / 1000.0 if unit.startswith("ns") else (val * 1000.0 if unit.startswith("ms") else val)
        name = r["Kernel Name"].split("(")[0]
        per.setdefault(name, []).append(us)
    total = sum(sum(v) for v in per.values())
    lines = [f"# {path.name}: every launch of `python bench.py --workload ... --steps 3 --warmup 3` with its device time",
             "# (ncu --metrics gpu__time_duration.sum --clock-control none; serialised, compare SHARES)",
             f"# launches: {sum(len(v) for v in per.values())}, total {total:.1f} us", "",
             f"{'kernel':60s} {'launches':>8s} {'avg us':>10s} {'sum us':>12s} {'share':>7s}"]
    for name, vals in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"{name[:60]:60s} {len(vals):8d} {sum(vals) / len(vals):10.2f} {sum(vals):12.1f} {100 * sum(vals) / total:6.1f}%")
    dest.write_text("\n".join(lines) + "\n")
    print("wrote", dest)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    for rep in sorted(OUT.glob(f"prof_*_{tag}.ncu-rep")):
        full_report(rep, ROOT / "profiles" / f"{tag}_{rep.stem.replace('prof_', '').replace('_' + tag, '')}_full.txt")
    for lst in sorted(OUT.glob(f"launches_*_{tag}.csv")):
        launch_list(lst, ROOT / "profiles" / f"{tag}_{lst.stem.replace('_' + tag, '')}.txt")


if __name__ == "__main__":
    main()
