#!/bin/bash
# Run under gpurun (one GPU): launch list of the bench + full ncu captures of the dominant kernels.
# usage: tools/profile.sh <tag>
# Numbers printed by a run under ncu are never bench values; the bench lines come from the plain runs at the end.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
# 1. every launch with its device time (cold cache, serialised: compare SHARES).  The pass is 92 launches; the
#    library's create-time eager pass comes first, then bench warm-ups and steps.
ncu --metrics gpu__time_duration.sum --clock-control none -s 96 -c 480 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/${TAG}_launches.log 2>&1
# 2. full captures.  k_resblock_tc: 21 launches per pass (layer1 x3, refine0/1/2 x6); skip the create pass and land on the
#    full-resolution refinement blocks of the first bench pass.  k_conv_stream: 58 launches per pass (firstconv.1/.2,
#    layer1.0.downsample, layer2 x33, layer3 x7, layer4 x6, lastconv.1, head.filter.1-4, conv3d_alone, conv_out x3): skip the
#    create pass + 6 and capture layer2.1.conv_a / conv_b (the shape of 30 launches), then head.filter.1 (3-D).
ncu --set full --clock-control none --import-source on -k regex:k_resblock_tc -s 36 -c 2 -f -o gpurun_out/${TAG}_resblock \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_resblock.log 2>&1
ncu -i gpurun_out/${TAG}_resblock.ncu-rep --page raw --csv > gpurun_out/${TAG}_resblock_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_conv_stream -s 64 -c 2 -f -o gpurun_out/${TAG}_stream \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_stream.log 2>&1
ncu -i gpurun_out/${TAG}_stream.ncu-rep --page raw --csv > gpurun_out/${TAG}_stream_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_conv_stream -s 108 -c 1 -f -o gpurun_out/${TAG}_stream3d \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_stream3d.log 2>&1
ncu -i gpurun_out/${TAG}_stream3d.ncu-rep --page raw --csv > gpurun_out/${TAG}_stream3d_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_costvol -s 1 -c 1 -f -o gpurun_out/${TAG}_costvol \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_costvol.log 2>&1
ncu -i gpurun_out/${TAG}_costvol.ncu-rep --page raw --csv > gpurun_out/${TAG}_costvol_raw.csv 2>/dev/null
# 3. plain runs: per-op CUDA-event table, bench lines
python tools/opprof.py --precision tc > gpurun_out/${TAG}_opprof.txt 2>&1
SNB_TC_PROF=1 python tools/opprof.py --precision tc --reps 1 2>&1 | grep -E "rbprof|csprof|tcprof" > gpurun_out/${TAG}_roleprof.txt
python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_reference_bench.json 2>> gpurun_out/${TAG}_bench.err
ls -la gpurun_out/ | tail -30
