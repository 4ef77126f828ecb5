for o in "enc_persist=0" "enc_pchunk=8" "enc_pchunk=12" "enc_pchunk=16" "enc_pchunk=24" "enc_pchunk=32"; do
  echo "== $o"; timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-beam --opt $o 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])"
done
for o in "enc_persist=0" "enc_pchunk=8" "enc_pchunk=16"; do echo "== stage $o"; timeout -s KILL 100 python tools/profile_step.py --precision tf32 --steps 3 --opt stage_timing=1 --opt $o 2>&1 | tail -12; done
