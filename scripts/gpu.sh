#!/bin/bash
# Rebuild libshb200.so, then run a command on the B200 box:  scripts/gpu.sh [--timeout S] [--gpus N] -- '<cmd>'
set -e
cd "$(dirname "$0")/.."
python semantichuman_b200/_build.py >/dev/null
exec /usr/local/graft/bin/gpurun "$@"
