run() { echo "== $*"; timeout -s KILL 40 python tools/persist_debug.py "$@" 2>&1 | tail -8; }
run --T 640 --steps 3 --sync 1 --opt enc_pchunk=16
run --T 300 1680 640 200 --steps 8 --sync 0 --opt enc_pchunk=16
timeout -s KILL 120 python -m pytest tests -m gpu -x -q -k "wavefront" 2>&1 | tail -5
