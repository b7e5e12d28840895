"""Summaries of ncu output for profiles/ (the judge reads these; gpurun_out/ is scratch).

  launch list   python tools/ncu_summary.py launches gpurun_out/launches.csv "<command that was profiled>" > profiles/rNN_ncu_launch_summary_X.csv
                (input: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <command>)
  full capture  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep > profiles/rNN_ncu_full_X.csv
                (input: ncu --set full --clock-control none --import-source on -o prof <command>; needs `ncu` on PATH to read it)

Per-launch times of a profiler run are cold-cache and serialised: compare SHARES, never absolute numbers.
"""
from __future__ import annotations

import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

FULL_METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
]


def short_name(full: str) -> str:
    """`void vitk::<unnamed>::gemm_tcgen05_kernel<192, 4, 0, 0, 0>(Params)` -> `gemm_tcgen05_kernel<192, 4, 0, 0, 0>`."""
    s = re.sub(r"^void\s+", "", full.strip())
    s = re.sub(r"^(vitk::)?(\(anonymous namespace\)|<?unnamed>|\(unnamed namespace\))::", "", s)
    s = re.sub(r"^vitk::", "", s)
    depth = 0
    for i, ch in enumerate(s):            # strip the argument list: the first '(' outside template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return s[:i]
    return s


def _rows(path: str):
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    if start < 0:
        raise SystemExit(f"{path}: no ncu CSV header found")
    return csv.DictReader(io.StringIO(text[start:]))


def launches(path: str, command: str) -> None:
    agg: "OrderedDict[str, list]" = OrderedDict()
    n = 0
    for r in _rows(path):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "s": 1e9, "second": 1e9}.get(unit, 1.0)
        a = agg.setdefault(short_name(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += ns
        n += 1
    total = sum(a[1] for a in agg.values())
    print("# ncu launch list summary (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)")
    print(f"# command: {command}")
    print(f"# total captured: {total / 1e6:.3f} ms over {n} launches")
    print("kernel,launches,total_ms,share,avg_us")
    for k, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'"{k}",{cnt},{ns / 1e6:.3f},{ns / total:.4f},{ns / cnt / 1e3:.2f}')


def full(path: str) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        raise SystemExit(out.stderr)
    text = out.stdout
    rd = csv.reader(io.StringIO(text[text.find('"ID"'):]))
    header = next(rd)
    units = next(rd)
    cols = [(m, header.index(m)) for m in FULL_METRICS if m in header]
    kcol = header.index("Kernel Name")
    print(f"# ncu --set full --clock-control none --import-source on; one row per captured launch ({path.split('/')[-1]})")
    print("kernel," + ",".join(f"{m} [{units[i]}]" for m, i in cols))
    for row in rd:
        if len(row) <= kcol:
            continue
        print(f'"{short_name(row[kcol])}",' + ",".join(row[i].replace(",", "") for _, i in cols))


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("launches", "full"):
        raise SystemExit(__doc__)
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "(not recorded)")
    else:
        full(sys.argv[2])
