#!/bin/bash
# Profiling recipe (run under gpurun, 1 GPU). Usage: bash tools_profile.sh <tag>
#   per workload: launch list with DRAM bytes of one bench step (gpurun_out/<tag>_launches_<w>.csv + .json, read here by
#   tools_ncu_traffic.py) and one ncu --set full capture of the hot kernels (gpurun_out/<tag>_prof_<w>.ncu-rep).
TAG=${1:-prof}
mkdir -p gpurun_out
for spec in "readme 16000000 tailwalk|nl_count" "syslog200 16000000 tailwalk|dfawalk|linewalk" "weblog 16000000 tailwalk|linewalk|capwalk" "utf16mix 16000000 tailwalk|linewalk|capwalk" "simple 16000000 tailwalk"; do
    set -- $spec
    W=$1; LINES=$2; KREGEX=$3
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_$W.csv \
        python bench.py --workload $W --steps 1 --warmup 3 --lines-per-gpu $LINES --skip-e2e --skip-cpu --configs "" > gpurun_out/${TAG}_launches_$W.json 2> gpurun_out/${TAG}_launches_$W.err
    if [ "$W" != "simple" ] && [ "$W" != "utf16mix" ]; then
        ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s 6 -c 4 -f -o gpurun_out/${TAG}_prof_$W \
            python bench.py --workload $W --steps 1 --warmup 3 --lines-per-gpu $LINES --skip-e2e --skip-cpu --configs "" > gpurun_out/${TAG}_prof_$W.log 2>&1
    fi
done
ls -la gpurun_out/ | tail -20
