"""Top stall sites per kernel from an ncu source-page CSV:
   ncu -i X.ncu-rep --page source --csv > f.csv;  python tools/ncu_top.py f.csv [N] [kernel-substring] [section-index]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
pat = sys.argv[3] if len(sys.argv) > 3 else ""
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
secs = [(rows[s][1], rows[s + 1], rows[s + 2:(starts[k + 1] if k + 1 < len(starts) else len(rows))]) for k, s in enumerate(starts)]
secs = [s for s in secs if pat in s[0]]
print(f"{len(secs)} matching kernel sections")
name, hdr, data = secs[which]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in data if len(r) == len(hdr)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("kernel:", name[:160], " total samples", tot, " instructions", len(data))
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print("stall totals:", {k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:n]
for i in sorted(top):
    r = data[i]
    st = {s[6:]: int(r[ix[s]] or 0) for s in stalls if int(r[ix[s]] or 0) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{i:5d} {int(r[ix['# Samples']]):6d} exec={r[ix['Instructions Executed']]:>8s} {r[ix['Source']].strip()[:80]:80s} {st}")
