#!/bin/bash
# Profiling recipe (run under gpurun, 1 GPU): launch list of one bench step + full ncu capture of the hot kernels.
# Usage: bash tools_profile.sh <tag> [lines] [kernel-regex]
#   -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_prof.ncu-rep
set -x
TAG=${1:-prof}
LINES=${2:-8000000}
KREGEX=${3:-'onepass|fused|dfa_|tdfa_|nl_'}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --lines-per-gpu $LINES --skip-e2e --skip-cpu > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s 3 -c 2 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 1 --warmup 3 --lines-per-gpu $LINES --skip-e2e --skip-cpu > gpurun_out/${TAG}_prof.log 2>&1
ls -la gpurun_out/
