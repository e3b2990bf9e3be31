#!/bin/bash
# Run under gpurun: launch list of one bench step + one full ncu capture of the top kernel.
# usage: tools/profile.sh <tag> <precision fp32|tc> <kernel-regex>
set -u
TAG=${1:-r01}; PREC=${2:-fp32}; KRE=${3:-k_conv_direct}
mkdir -p gpurun_out
# launches per step ~115; skip create warm-up + 3 bench warm-ups, then list two steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 460 -c 240 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --precision $PREC --no-cpu-baseline \
    > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 150 -c 3 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 3 --warmup 3 --precision $PREC --no-cpu-baseline > gpurun_out/${TAG}_prof.log 2>&1
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
ls -la gpurun_out/
