timeout -s KILL 120 python -m pytest tests -m gpu -x -q -k "wavefront or bucket" 2>&1 | tail -3
st() { echo "== stage $*"; timeout -s KILL 40 python tools/profile_step.py --precision tf32 --steps 6 --opt stage_timing=1 "$@" 2>&1 | grep stages | tail -5 | cut -c1-200; }
st --opt enc_persist=3 --opt enc_pchunk=8
st --opt enc_persist=3 --opt enc_pchunk=8 --opt enc_side_ctas=8
st --opt enc_persist=3 --opt enc_pchunk=8 --opt enc_gemm_ctas_bwd=8
for o in "enc_persist=1" "enc_persist=3"; do
  echo "== $o"; timeout -s KILL 60 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-beam --opt $o --opt enc_pchunk=8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])"
done
