g="python tools/gemm_roofline_shapes.py"
timeout -s KILL 60 $g || exit 1
timeout -s KILL 200 ncu -f --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 3 -c 1 -o /tmp/cap_a $g > gpurun_out/ncu_f_l0proj_v28.log 2>&1; echo rc=$?
ncu -i /tmp/cap_a.ncu-rep --page raw --csv > gpurun_out/full_l0proj_v28.csv 2>/dev/null
timeout -s KILL 200 ncu -f --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o /tmp/cap_b $g > gpurun_out/ncu_f_cnndx_v28.log 2>&1; echo rc=$?
ncu -i /tmp/cap_b.ncu-rep --page raw --csv > gpurun_out/full_cnndx_v28.csv 2>/dev/null
python tools/ncu_extract.py "grouped 2-CTA tcgen05 GEMM, the two encoder layer-0 input projections 2 x (M5120 N1024 K1536) (NT): the bench.py roofline kernel=gpurun_out/full_l0proj_v28.csv" "1-CTA persistent tcgen05 GEMM, CNN_1 data gradient as a transposed convolution, even rows: M15744 N128 K2560 (NN, overlapping-rows A)=gpurun_out/full_cnndx_v28.csv" > gpurun_out/extract_v28.txt
cut -c1-600 gpurun_out/extract_v28.txt
