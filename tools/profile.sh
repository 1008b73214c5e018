#!/bin/bash
# Profiling recipe of this repo (run on a B200 through gpurun; see /opt/skills/guides/B200_PROFILING.md).
#   bash tools/profile.sh <tag> [launches|full|both]
# Profiles the benchmark's OWN configuration (10 M sites / 22 chromosomes, -s 1024; `--profile` keeps only the
# device-resident timed steps).  Writes
#   gpurun_out/<tag>_launches.csv   every launch with its device time (ncu, serialised, cold cache)
#   gpurun_out/<tag>_prof.ncu-rep   ncu --set full (+ source) of the first two scan_kernel launches of the
#                                   timed step: chromosome 1 (846 k sites) and chromosome 2
set -u
tag=${1:-r2}
what=${2:-both}
cmd="python bench.py --sites ${SITES:-10000000} --steps 1 --warmup 3 --profile ${BENCH_ARGS:-}"
mkdir -p gpurun_out
if [ "$what" != full ]; then
$cmd > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu_launches.log 2>&1
fi
if [ "$what" != launches ]; then
$cmd > gpurun_out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 66 -c 2 \
    -o gpurun_out/${tag}_prof $cmd > gpurun_out/${tag}_ncu_full.log 2>&1
fi
for f in gpurun_out/${tag}_plain*.log; do tail -n 1 "$f" | cut -c1-300; done
