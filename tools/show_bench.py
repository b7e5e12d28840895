import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f"{f}: value {d['value']:.0f} img/s  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.0f}  launches/step {d.get('launches_per_step')}  clocks {d['clocks']}")
    r = d.get("roofline") or {}
    print("  roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k not in ('algorithmic_per_launch', 'peak_source')})
    if d.get("cpu_baseline"): print("  cpu", d["cpu_baseline"])
    for k, v in (d.get("kernels") or {}).items():
        print(f"   {k:24s} n={v['launches_per_step']:3d} ms={v['ms_per_step']:7.3f} share={v['share']:.3f} tflops={v.get('tflops', 0):7.1f} gbs={v['gbs']:7.0f}")
