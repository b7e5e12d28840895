export VITK_LIB=$PWD/thyroid-vit-cnn-comparison_b200/libvitk_dbg.so
for k in 0 14 60; do echo "== knobs $k"; VITK_GEMM_KNOBS=$k python tools/gemm_scan.py 576 192; VITK_GEMM_KNOBS=$k python tools/gemm_scan.py 192 768;  VITK_GEMM_KNOBS=$k python tools/gemm_scan.py 768 768; done
