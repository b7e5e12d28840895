"""Run every GPU kernel test in its own subprocess (a faulting kernel must not poison the CUDA
context of the others) and write a compact report to gpurun_out/probe_report.txt."""
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)


def main():
    test_file = sys.argv[1] if len(sys.argv) > 1 else "tests/test_kernels_gpu.py"
    r = subprocess.run([sys.executable, "-m", "pytest", test_file, "-m", "gpu", "--collect-only", "-q"],
                       cwd=ROOT, capture_output=True, text=True)
    ids = [l.strip() for l in r.stdout.splitlines() if "::" in l]
    # group by test function so that each subprocess stays short
    groups = {}
    for i in ids:
        groups.setdefault(i.split("[")[0], []).append(i)
    lines = []
    for name, members in groups.items():
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "--no-header", "-p", "no:cacheprovider",
                                "--tb=short", *members], cwd=ROOT, capture_output=True, text=True, timeout=300)
            status = "PASS" if p.returncode == 0 else f"FAIL rc={p.returncode}"
            tail = (p.stdout + p.stderr)[-3000:] if p.returncode != 0 else ""
        except subprocess.TimeoutExpired:
            status, tail = "TIMEOUT", ""
        lines.append(f"{status:12s} {name}  ({time.time() - t0:.1f}s)")
        if tail:
            lines.append(tail)
        print(lines[-2 if tail else -1], flush=True)
    (OUT / "probe_report.txt").write_text("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
