O=gpurun_out/r01m; mkdir -p $O
timeout 200 ncu --set full --import-source on --clock-control none -k 'regex:attn_bwd_tc|attn_fwd_tc' --launch-skip 1 -c 2 -f -o $O/attn python tools/kbench.py --model deit_tiny --iters 1 --warm 0 --only attn_fwd,attn_bwd > $O/ncu_attn.log 2>&1
ncu -i $O/attn.ncu-rep --page source --csv > $O/attn_src.csv 2>/dev/null
python tools/ncu_top.py $O/attn_src.csv 28 attn_bwd > $O/stalls_attn_bwd.txt 2>&1
python tools/ncu_top.py $O/attn_src.csv 22 attn_fwd > $O/stalls_attn_fwd.txt 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k 'regex:gemm_tcgen05' -c 2 -f -o $O/gemm python tools/kbench.py --model deit_tiny --iters 1 --warm 0 --only gemm_qkv,gemm_fc2 > $O/ncu_gemm.log 2>&1
ncu -i $O/gemm.ncu-rep --page source --csv > $O/gemm_src.csv 2>/dev/null
python tools/ncu_top.py $O/gemm_src.csv 28 gemm 0 > $O/stalls_gemm_qkv.txt 2>&1
python tools/ncu_top.py $O/gemm_src.csv 28 gemm 1 > $O/stalls_gemm_fc2.txt 2>&1
rm -f $O/*.ncu-rep $O/*_src.csv
du -sh $O; head -5 $O/stalls_attn_bwd.txt
