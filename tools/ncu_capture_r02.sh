#!/bin/bash
# Round-2 ncu evidence: launch list of one training step (serialisable structure) + --set full captures, exported to CSV on the box.
# Usage: bash tools/ncu_capture_r02.sh <tag> [what]
tag=${1:-r02}
what=${2:-all}
export AST_NO_COOP=1      # ncu rejects cooperative + cluster launches; co-residency holds anyway (128 CTAs, idle GPU)
step="python tools/profile_step.py --precision tf32 --steps 1 --warmup 1"
mkdir -p gpurun_out
cap() {  # name regex skip count cmd...
    local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
    ncu -f --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o /tmp/cap_$name "$@" > gpurun_out/ncu_f_${name}_$tag.log 2>&1
    echo "capture $name rc=$?"
    ncu -i /tmp/cap_$name.ncu-rep --page raw --csv > gpurun_out/full_${name}_$tag.csv 2>/dev/null
    rm -f /tmp/cap_$name.ncu-rep
}
if [ "$what" = all ] || [ "$what" = list ]; then
    $step > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
    ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$tag.csv $step > gpurun_out/ncu_l_$tag.log 2>&1
    echo "launch list rc=$?"
fi
if [ "$what" = all ] || [ "$what" = dec ]; then cap dec "dec_seq2" 2 2 $step; fi
if [ "$what" = all ] || [ "$what" = lstm ]; then cap lstm "lstm_seq_(fwd|bwd)_tc" 50 4 $step; fi
if [ "$what" = all ] || [ "$what" = mem ]; then cap mem "bn_bwd|bn_stats|bn_relu_to_rnn|colsum4|softmax_ce_all|opt_amsgrad|fill_sentinel" 11 22 $step; fi
if [ "$what" = all ] || [ "$what" = pack ]; then cap pack "pack_cmvn" 2 1 python tools/ncu_drive_pack.py; fi
if [ "$what" = all ] || [ "$what" = beam ]; then cap beam "beam_topk_batch|beam_prune_batch|beam_gather_batch|attn_dot_grouped|attn_ctx_grouped|gemm_tc_kernel" 40 24 python tools/ncu_drive_beam.py; fi
du -sh gpurun_out
