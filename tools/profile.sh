#!/bin/bash
# Profiling recipe of this repo (run on a B200 through gpurun; see /opt/skills/guides/B200_PROFILING.md).
#   bash tools/profile.sh <tag>
# writes gpurun_out/<tag>_launches.csv (every launch with its device time) and
# gpurun_out/<tag>_prof.ncu-rep (ncu --set full of two scan_kernel launches of the timed step).
set -u
tag=${1:-r1}
cmd="python bench.py --sites 1000000 --steps 1 --warmup 3 --no-cpu ${BENCH_ARGS:-}"
mkdir -p gpurun_out
$cmd > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu_launches.log 2>&1
$cmd > gpurun_out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 66 -c 2 \
    -o gpurun_out/${tag}_prof $cmd > gpurun_out/${tag}_ncu_full.log 2>&1
tail -1 gpurun_out/${tag}_plain.log | cut -c1-300
