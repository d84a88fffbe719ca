#!/bin/bash
# launch list (per-kernel device time) of a short bench run; run under gpurun.  $1 = tag
set -e
TAG=${1:-r1}
python bench.py --steps 2 --warmup 3 --no-cpu --no-c5 --no-extra > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-c5 --no-extra > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log
