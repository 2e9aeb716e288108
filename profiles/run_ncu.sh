#!/bin/bash
# profiles/run_ncu.sh <tag> -- run under gpurun; writes gpurun_out/<tag>_launches.csv and gpurun_out/<tag>_top.ncu-rep
# (numbers printed by the profiled runs are NOT bench values)
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
# 1) every launch with its device time: a short steady-ish run (few blocks, one of them with a single sync)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-compress-e2e > gpurun_out/${TAG}_launches.log 2>&1
# 2) the top kernel, full set, 3 launches
ncu --set full --clock-control none --import-source on -k regex:${2:-k_walk} -s 20 -c 3 -o gpurun_out/${TAG}_top \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-compress-e2e > gpurun_out/${TAG}_top.log 2>&1
ls -la gpurun_out
