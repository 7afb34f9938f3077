#!/bin/bash
# Profiling recipe (run under gpurun, 1 GPU): launch lists of bench steps + full ncu captures of the hot kernels.
# Usage: bash tools_profile.sh <tag>
#   -> gpurun_out/<tag>_launches_<workload>.csv, gpurun_out/<tag>_prof_<workload>.ncu-rep
TAG=${1:-prof}
mkdir -p gpurun_out
for spec in "readme 8000000 chunkwalk" "syslog200 16000000 linewalk|capwalk|nl_" "weblog 16000000 linewalk|capwalk|nl_"; do
    set -- $spec
    W=$1; LINES=$2; KREGEX=$3
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_$W.csv \
        python bench.py --workload $W --steps 2 --warmup 3 --lines-per-gpu $LINES --skip-e2e --skip-cpu > gpurun_out/${TAG}_launches_$W.log 2>&1
    ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s 6 -c 4 -f -o gpurun_out/${TAG}_prof_$W \
        python bench.py --workload $W --steps 1 --warmup 3 --lines-per-gpu $LINES --skip-e2e --skip-cpu > gpurun_out/${TAG}_prof_$W.log 2>&1
done
ls -la gpurun_out/
