"""profiles/rNN_ncu_full_<model>.csv (tools/ncu_summary.py full) -> profiles/rNN_ncu_traffic_<model>.json, the per-launch DRAM
traffic table bench.py attaches to its roofline entries.  The capture runs tools/kbench.py --only <cases> --iters 1 --warm 0,
one launch per case, so row i of the CSV is case i of the --only list.

    python tools/ncu_traffic.py profiles/r02_ncu_full_deit_tiny.csv ln_fwd,ln_bwd,... > profiles/r02_ncu_traffic_deit_tiny.json
"""
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "%": 1.0}


def main():
    path, cases = sys.argv[1], sys.argv[2].split(",")
    lines = [l for l in open(path) if not l.startswith("#")]
    rd = csv.reader(lines)
    header = next(rd)
    cols = {}
    for i, h in enumerate(header):
        m = re.match(r"(.*) \[(.*)\]", h)
        cols[m.group(1) if m else h] = (i, m.group(2) if m else "")
    rows = list(rd)
    if len(rows) != len(cases):
        raise SystemExit(f"{path}: {len(rows)} captured launches for {len(cases)} cases")
    out = {}

    def val(row, name):
        i, unit = cols[name]
        return float(row[i]) * UNIT.get(unit, 1.0)

    for case, row in zip(cases, rows):
        out[case] = {
            "kernel": row[0],
            "duration_us": val(row, "gpu__time_duration.sum"),
            "dram_bytes": val(row, "dram__bytes_read.sum") + val(row, "dram__bytes_write.sum"),
            "dram_pct": val(row, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "tensor_pipe_pct": val(row, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "sm_throughput_pct": val(row, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
