#!/bin/bash
# tools/e2e_multi.sh <n_gpus>: the sweep of tools/e2e_multi.py (raw copies, default call, chunk / slot variants)
n=${1:-4}
run() {  # run <label> <extra args> (environment from the caller)
    T0=$(python -c "import time; print(time.time() + ${3:-30})")
    echo "== $1"
    for i in $(seq 0 $((n - 1))); do python tools/e2e_multi.py --gpu $i --t0 $T0 $2 & done
    wait
}
python -c "import torch" # page the image in
run "raw copies (305 MB in, 50 MB out)" --raw 45
run "default (chunk 65536, 4 slots)" ""
TREAT_B200_CHUNK=131072 run "chunk 131072" ""
TREAT_B200_CHUNK=262144 run "chunk 262144" ""
TREAT_B200_SLOTS=2 run "2 slots" ""
TREAT_B200_CHUNK=32768 run "chunk 32768" ""
