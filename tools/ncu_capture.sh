#!/bin/bash
# ncu evidence for one training step (B32 x T640 x L24, TF32 mode): launch list + --set full captures of the dominant
# kernels, exported to CSV on the box (the .ncu-rep files exceed the 64 MiB return limit and are deleted).
# Usage: bash tools/ncu_capture.sh <tag>
tag=${1:-x}
export AST_NO_COOP=1      # ncu rejects cooperative + cluster launches; co-residency holds anyway (128 CTAs, idle GPU)
cmd="python tools/profile_step.py --precision tf32 --steps 1 --warmup 1"
mkdir -p gpurun_out
$cmd > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv $cmd > gpurun_out/ncu_l_$tag.log 2>&1
echo "launch list rc=$?"
cap() {  # name regex skip count
    ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/cap_$1 $cmd > gpurun_out/ncu_f_$1_$tag.log 2>&1
    echo "capture $1 rc=$?"
    ncu -i /tmp/cap_$1.ncu-rep --page raw --csv > gpurun_out/full_$1_$tag.csv 2>/dev/null
    ncu -i /tmp/cap_$1.ncu-rep --page details --csv > gpurun_out/details_$1_$tag.csv 2>/dev/null
    ls -la /tmp/cap_$1.ncu-rep
}
what=${2:-all}
if [ "$what" = all ] || [ "$what" = dec ]; then cap dec "dec_seq2" 2 2; fi
if [ "$what" = all ] || [ "$what" = lstm ]; then cap lstm "lstm_seq_(fwd|bwd)_tc" 30 4; fi
if [ "$what" = all ] || [ "$what" = gemm ]; then
    cmd="python tools/profile_step.py --precision tf32 --steps 1 --warmup 0"     # launches 2,3 = the L0 input projections
    cap gemm "gemm_tc_kernel" 0 6
    cmd="python tools/profile_step.py --precision tf32 --steps 1 --warmup 1"
fi
if [ "$what" = all ] || [ "$what" = mem ]; then cap mem "split_tf32|opt_amsgrad|bn_relu_to_rnn|bn_bwd_apply" 8 6; fi
du -sh gpurun_out
