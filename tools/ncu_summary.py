#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small text files for profiles/.

    python tools/ncu_summary.py report <file.ncu-rep> [...]      key metrics of every kernel in a `--set full` report
    python tools/ncu_summary.py launches <launches.csv>          per-kernel totals of a gpu__time_duration launch list
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_op_dmma.sum", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_lsu.sum",
]


def report(path):
    if path.endswith(".csv"):     # already exported on the GPU box (the .ncu-rep files are too large to bring back)
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"## kernel {d.get('Kernel Name')}  (id {d.get('ID')})")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"{k:78s} {d[k]:>16s} {u[k]}")
        stalls = [(float(d[k].replace(',', '')), k) for k in hdr
                  if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and d[k]]
        if not stalls:
            stalls = [(float(d[k].replace(',', '')), k) for k in hdr
                      if "issue_stalled" in k and k.endswith(".pct") and d[k]]
        for v, k in sorted(stalls, reverse=True)[:8]:
            print(f"{k:78s} {v:16.3f}")
        try:
            t = float(d["gpu__time_duration.sum"].replace(",", ""))
            tu = u["gpu__time_duration.sum"]
            scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu, {"nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}.get(tu, 1e-9))
            def b(k):
                v = float(d[k].replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]
            tot = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
            print(f"{'derived: dram bytes per launch':78s} {tot:16.0f} byte")
            print(f"{'derived: dram GB/s under ncu (cold, serialised)':78s} {tot / (t * scale) / 1e9:16.1f} GB/s")
        except Exception as e:  # noqa: BLE001
            print("derived: n/a", e)
        print()


def launches(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    tot = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        name = r["Kernel Name"].split("(")[0]
        tot[name][0] += 1
        tot[name][1] += v
    allus = sum(v[1] for v in tot.values())
    print(f"# {path}: kernel totals (ncu gpu__time_duration.sum, --clock-control none; cold-cache, serialised launches)")
    print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:70]:70s} {n:8d} {us:12.1f} {us / n:10.2f} {us / allus:7.1%}")


if __name__ == "__main__":
    mode = sys.argv[1]
    for p in sys.argv[2:]:
        report(p) if mode == "report" else launches(p)
