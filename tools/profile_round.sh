#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun): plain benches first (their numbers are the reported ones),
# then the ncu launch lists of the same bench command and one `--set full` capture per model of the isolated hot kernels
# (one launch per kernel class, L2 flushed before it).  Only summaries travel back (the .ncu-rep files exceed the pull limit).
# Usage: bash tools/profile_round.sh <tag> [stages]     (writes gpurun_out/<tag>/; stages default "bench kbench launches full")
set -u
TAG=${1:-r01f}
STAGES=${2:-"bench kbench launches full"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
has() { [[ " $STAGES " == *" $1 "* ]]; }
if has bench; then
  timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err          # DeiT-tiny + models.vit_base
  timeout 300 python bench.py --mode distill --steps 20 --warmup 5 > $OUT/bench_distill.json 2> $OUT/bench_distill.err
  timeout 300 python bench.py --mode ensemble --steps 10 --warmup 3 > $OUT/bench_ensemble.json 2> $OUT/bench_ensemble.err
  timeout 300 python bench.py --dtype bf16 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_bf16.json 2> $OUT/bench_bf16.err
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference_arm.json 2> $OUT/bench_reference_arm.err
fi
if has newfull; then
  # one --set full capture of the round-2 kernels (isolated, one launch each)
  ONLY=tiles_to_patches,attn_probs,rollout_row,bottleneck_b1,bottleneck_b3
  timeout 300 ncu --set full --clock-control none -k 'regex:tiles_to_patches|attn_probs|rollout_row|dense_bottleneck' -c 5 -f -o $OUT/full_new \
      python tools/kbench.py --model deit_tiny --iters 1 --warm 0 --only $ONLY > $OUT/ncu_full_new.log 2>&1
  python tools/ncu_summary.py full $OUT/full_new.ncu-rep > $OUT/ncu_full_new_kernels.csv 2> $OUT/ncu_full_new.err
  rm -f $OUT/full_new.ncu-rep
fi
if has distill_launches; then
  CMD="python bench.py --mode distill --steps 1 --warmup 1 --no-graph --no-cpu-baseline"
  # the timed region only (bench.py brackets it with the NVTX range "timed"): the teacher's set-up casts ~1500 tensors first
  timeout 300 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_distill.csv $CMD > $OUT/ncu_launches_distill.log 2>&1
  python tools/ncu_summary.py launches $OUT/launches_distill.csv "ncu --nvtx --nvtx-include timed/ --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv $CMD" > $OUT/ncu_launch_summary_distill.csv
  rm -f $OUT/launches_distill.csv
fi
if has kbench; then
  timeout 120 python tools/kbench.py --model deit_tiny --json $OUT/kb_deit_tiny.json > $OUT/kb_deit_tiny.log 2>&1
  timeout 120 python tools/kbench.py --model vit_base --json $OUT/kb_vit_base.json > $OUT/kb_vit_base.log 2>&1
fi
for M in deit_tiny vit_base; do
  if has launches; then
    CMD="python bench.py --model $M --steps 1 --warmup 1 --no-graph --no-cpu-baseline"   # --model: this model alone (no sub-record)
    date +%s > $OUT/t0_launches_$M
    timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_$M.csv $CMD > $OUT/ncu_launches_$M.log 2>&1
    python tools/ncu_summary.py launches $OUT/launches_$M.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv $CMD" > $OUT/ncu_launch_summary_$M.csv
    rm -f $OUT/launches_$M.csv
    date +%s > $OUT/t1_launches_$M
  fi
  if has full; then
    ONLY=ln_fwd,ln_bwd,gemm_qkv,gemm_gelu,gemm_fc2,dgrad_fc1,wgrad_fc1,attn_fwd,attn_bwd
    timeout 300 ncu --set full --clock-control none -k 'regex:gemm_tcgen05|attn_|ln_' --launch-skip 2 -c 9 -f -o $OUT/full_$M \
        python tools/kbench.py --model $M --iters 1 --warm 0 --only $ONLY > $OUT/ncu_full_$M.log 2>&1
    python tools/ncu_summary.py full $OUT/full_$M.ncu-rep > $OUT/ncu_full_$M.csv 2> $OUT/ncu_full_$M.err
    ls -la $OUT/full_$M.ncu-rep >> $OUT/sizes.txt
    rm -f $OUT/full_$M.ncu-rep
    date +%s > $OUT/t1_full_$M
  fi
done
du -sh $OUT gpurun_out
