timeout -s KILL 200 python -m pytest tests -m gpu -x -q -k "wavefront or bucket or full_size or tf32 or gemm" 2>&1 | tail -3
for o in "tc2=0" "tc2=1"; do
  echo "== $o"; timeout -s KILL 60 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-beam --opt $o 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])"
done
