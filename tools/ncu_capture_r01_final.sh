#!/bin/bash
# Round-1 final ncu evidence.  (1) launch list of one training step; ncu serialises kernels, so the kernels that WAIT for one another
# (persistent encoder wavefront, cooperative decoder) are switched to their per-chunk / non-cooperative launches by AST_NO_COOP=1 -
# same kernels, same arithmetic, different launch structure.  (2) --set full captures of the 2-CTA tcgen05 GEMM (layer-0 data gradient
# shape and a large square-ish shape) and of the 1-CTA kernel at the CNN_1 data gradient shape (the bench.py roofline kernel).
tag=${1:-v26}
mkdir -p gpurun_out
export AST_NO_COOP=1
cmd="python tools/profile_step.py --precision tf32 --steps 1 --warmup 1"
timeout -s KILL 120 $cmd > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$tag.csv $cmd > gpurun_out/ncu_l_$tag.log 2>&1
echo "launch list rc=$?"
unset AST_NO_COOP
g="python tools/check_gemm_tc2.py"
timeout -s KILL 120 $g > gpurun_out/plain_gemm2_$tag.log 2>&1 || { echo "plain gemm run failed"; exit 1; }
cap() {  # name regex skip count
    timeout -s KILL 300 ncu -f --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/cap_$1 $g > gpurun_out/ncu_f_$1_$tag.log 2>&1
    echo "capture $1 rc=$?"
    ncu -i /tmp/cap_$1.ncu-rep --page raw --csv > gpurun_out/full_$1_$tag.csv 2>/dev/null
}
# check_gemm_tc2.py: 10 correctness launches of the 2-CTA kernel, then 8 timed launches per shape (2-CTA) interleaved with 8 of the 1-CTA kernel
cap tc2_l0dx "gemm_tc2_kernel" 21 1        # 5120 x 1536 x 1024 (NN)
cap tc2_big "gemm_tc2_kernel" 37 1         # 8192 x 4096 x 4096 (NT)
cap tc1_cnndx "gemm_tc_kernel" 19 1        # 1-CTA kernel, 15744 x 1152 x 512 (NN): launches 16..23 of that kernel
cap tc2_cnndx "gemm_tc2_kernel" 29 1       # 2-CTA kernel, same shape
python tools/ncu_extract.py "2-CTA tcgen05 GEMM, layer-0 data gradient shape M5120 N1536 K1024 (NN)=gpurun_out/full_tc2_l0dx_$tag.csv" \
   "2-CTA tcgen05 GEMM, M8192 N4096 K4096 (NT)=gpurun_out/full_tc2_big_$tag.csv" \
   "1-CTA persistent tcgen05 GEMM, CNN_1 data gradient M15744 N1152 K512 (NN): the bench.py roofline kernel=gpurun_out/full_tc1_cnndx_$tag.csv" \
   "2-CTA tcgen05 GEMM at the same CNN_1 data gradient shape=gpurun_out/full_tc2_cnndx_$tag.csv" > gpurun_out/extract_$tag.txt
cat gpurun_out/extract_$tag.txt | cut -c1-400
