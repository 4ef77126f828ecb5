#!/bin/bash
# One GPU round trip: GEMM check, GPU parity suite, bench (N=1), then the ncu launch list of one training step.
# Usage (from the repo root, on the GPU box): bash tools/gpu_cycle.sh <tag> [bench-steps]
tag=${1:-x}; steps=${2:-20}
mkdir -p gpurun_out
timeout 300 python tools/check_gemm_tc.py > gpurun_out/gemm_$tag.log 2>&1; echo "gemm rc=$?"; tail -8 gpurun_out/gemm_$tag.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
timeout 600 python bench.py --steps $steps --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$tag.json
timeout 300 python tools/profile_step.py --precision tf32 --steps 2 > gpurun_out/prof_$tag.log 2>&1; echo "prof rc=$?"; tail -2 gpurun_out/prof_$tag.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$tag.csv \
    python tools/profile_step.py --precision tf32 --steps 1 --warmup 1 > gpurun_out/ncu_$tag.log 2>&1; echo "ncu rc=$?"
