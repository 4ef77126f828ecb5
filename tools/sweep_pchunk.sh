#!/bin/bash
# A/B an engine option on the GPU box: parity subset, stage timing and the bucketed bench for each value.
# Usage: bash tools/sweep_pchunk.sh "enc_pchunk=4" "enc_pchunk=8" "enc_persist=0" ...   (every command under a kill timeout: a
# schedule that deadlocks must not hang the box)
timeout -s KILL 200 python -m pytest tests -m gpu -x -q -k "wavefront or bucket or full_size or tf32" 2>&1 | tail -2
for o in "$@"; do
  echo "== $o"
  timeout -s KILL 40 python tools/profile_step.py --precision tf32 --steps 5 --opt stage_timing=1 --opt $o 2>&1 | grep stages | tail -1 | cut -c1-220
  timeout -s KILL 60 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-beam --opt $o 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])"
done
