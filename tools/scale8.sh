#!/bin/bash
# Round-end multi-GPU bench lines on N GPUs of one box: train step (DeiT-tiny + ViT-B/16 sub-record), distillation step, ensemble.
N=$1; OUT=$2; mkdir -p $OUT
run() { name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus $N "$@" > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d.get("dp") and d["dp"]["buckets"], flush=True)
    if "models" in d:
        m = d["models"]["vit_base"]
        print("   vit_base", "value %.0f" % m["value"], "ms %.3f" % m["ms_per_step"], "e2e %.0f" % m["e2e"]["value"], m["clocks"], m["dp"]["buckets"], m["dp"]["launched_before_backward_ended"], flush=True)
except Exception as e:
    print("$name FAILED", e, open("$OUT/$name.err").read()[-800:], flush=True)
PY
}
run ce --steps 20 --warmup 5
run distill --mode distill --steps 20 --warmup 5
run ensemble --mode ensemble --steps 10 --warmup 3
