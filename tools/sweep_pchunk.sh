run() { echo "== $*"; timeout -s KILL 40 python tools/persist_debug.py "$@" 2>&1 | tail -14; }
run --T 640 --steps 3 --sync 1
run --T 300 1680 640 200 --steps 8 --sync 0
timeout -s KILL 120 python -m pytest tests -m gpu -x -q -k "wavefront or bucket" 2>&1 | tail -5
st() { echo "== stage $*"; timeout -s KILL 40 python tools/profile_step.py --precision tf32 --steps 3 --opt stage_timing=1 "$@" 2>&1 | tail -2; }
st --opt enc_persist=0
st --opt enc_pchunk=16
st --opt enc_pchunk=8
st --opt enc_pchunk=4
st --opt enc_pchunk=8 --opt enc_gemm_ctas=4
st --opt enc_pchunk=8 --opt enc_gemm_ctas=12
