#!/bin/bash
# Schedule sweep of the persistent encoder wavefront (bench.py step time per setting).  Usage: tools/sweep_sched.sh "VAR=val VAR=val" ...
for cfg in "$@"; do
  out=$(env $cfg timeout 200 python bench.py --no-cpu-baseline --no-sub --no-beam 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])")
  echo "$cfg -> $out"
done
