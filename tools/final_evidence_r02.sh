#!/bin/bash
# End-of-round evidence on one B200: GPU tests, the default bench line, recurrence probe / wavefront timeline / step timeline, ncu launch
# list and a --set full capture of the encoder recurrence kernels.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/final_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
timeout 120 python tools/enc_step_probe.py > gpurun_out/final_enc_step_probe.txt 2>&1
timeout 120 python tools/enc_timeline.py > gpurun_out/final_enc_timeline.txt 2>&1
timeout 200 python tools/kineto_step.py --timeline > gpurun_out/final_timeline.txt 2>&1
timeout 400 bash tools/ncu_capture_r02.sh r02c list
timeout 300 bash tools/ncu_capture_r02.sh r02c lstm
