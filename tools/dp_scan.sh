#!/bin/bash
# DP scaling scan on N GPUs of one box: DeiT-tiny train step with different bucket counts / NCCL CTA caps.
# usage: tools/dp_scan.sh N out_dir
N=$1; OUT=$2; mkdir -p $OUT
run() {  # name, extra args...
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus $N --steps 30 --warmup 5 --no-second-model --no-cpu-baseline "$@" > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "e2e_ms %.3f" % d["e2e"]["ms_per_step"],
          "buckets", d["dp"]["buckets"], "early", d["dp"]["launched_before_backward_ended"], "ctas", d["dp"]["nccl_max_ctas"], flush=True)
except Exception as e:
    print("$name FAILED", e, open("$OUT/$name.err").read()[-600:], flush=True)
PY
}
run b6 --min-buckets 6
run b1 --min-buckets 1
run b3 --min-buckets 3
run b6_c4 --min-buckets 6 --nccl-max-ctas 4
run b6_c8 --min-buckets 6 --nccl-max-ctas 8
run b2_c8 --min-buckets 2 --nccl-max-ctas 8
