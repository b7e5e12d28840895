export VITK_LIB=$PWD/thyroid-vit-cnn-comparison_b200/libvitk_dbg.so
for k in 0 1 2 12 14; do echo "== knobs $k"; VITK_GEMM_KNOBS=$k python tools/kbench.py --model deit_tiny --only gemm_qkv,gemm_gelu,gemm_fc2,dgrad_fc1,gemm_proj,wgrad_fc2 --iters 6 2>&1 | cut -c1-60; done
