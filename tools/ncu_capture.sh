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
    ncu -f --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/cap_$1 $cmd > gpurun_out/ncu_f_$1_$tag.log 2>&1
    echo "capture $1 rc=$?"
    ncu -i /tmp/cap_$1.ncu-rep --page raw --csv > gpurun_out/full_$1_$tag.csv 2>/dev/null
    ncu -i /tmp/cap_$1.ncu-rep --page details --csv > gpurun_out/details_$1_$tag.csv 2>/dev/null
    ls -la /tmp/cap_$1.ncu-rep
}
what=${2:-all}
if [ "$what" = all ] || [ "$what" = dec ]; then cap dec "dec_seq2" 2 2; fi
if [ "$what" = all ] || [ "$what" = lstm ]; then cap lstm "lstm_seq_(fwd|bwd)_tc" 50 4; fi
if [ "$what" = all ] || [ "$what" = gemm ]; then
    # the step's largest single-pass GEMM is the CNN_1 data gradient (NN form, M = B*F'*Rs, N = 1152, K = 512): find its index
    # among the gemm_tc launches of the launch list, capture it and the 3xTF32 forward convolution (first launch of the step)
    skip=$(python - "$tag" <<'PY'
import csv, sys
rows = [r for r in csv.DictReader(l for l in open(f"gpurun_out/launches_{sys.argv[1]}.csv") if l.startswith('"')) if r.get("Metric Name") == "gpu__time_duration.sum"]
g = [(i, r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for i, r in enumerate(r for r in rows if "gemm_tc_kernel" in r["Kernel Name"])]
half = len(g) // 2                      # second step (after the warm-up step)
nn = [(d, i) for i, k, d in g[half:] if "<0, 1, 0>" in k]
print(max(nn)[1] if nn else half)
PY
)
    echo "conv1 dx gemm index: $skip"
    cap gemm "gemm_tc_kernel" $skip 1
    mv gpurun_out/full_gemm_$tag.csv gpurun_out/full_gemmdx_$tag.csv; mv gpurun_out/details_gemm_$tag.csv gpurun_out/details_gemmdx_$tag.csv
    cmd="python tools/profile_step.py --precision tf32 --steps 1 --warmup 0"
    cap gemm "gemm_tc_kernel" 0 3
    cmd="python tools/profile_step.py --precision tf32 --steps 1 --warmup 1"
fi
if [ "$what" = all ] || [ "$what" = mem ]; then cap mem "split_tf32|opt_amsgrad|bn_relu_to_rnn|bn_bwd_rnn" 7 7; fi
du -sh gpurun_out
