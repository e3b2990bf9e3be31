#!/bin/bash
# Run under gpurun (one GPU): launch list of the bench + full ncu captures of the dominant kernels.
# usage: tools/profile.sh <tag>
# Numbers printed by a run under ncu are never bench values; the bench lines come from the plain runs at the end.
# Launch arithmetic (config 2, K = 3, round-2 pipeline): a pass is 85 launches, all inside the CUDA graph:
#   k_conv_stream 55 (firstconv.1/.2, layer2 x32, layer3 x6, layer4 x6, head.filter.0 as two halves, head.filter.1-4, conv_out x3),
#   k_resblock_tc 21 (layer1 x3, refine0/1/2 x6), k_refine_head 3, k_conv_tc 1 (lastconv.0), k_conv1x1 2,
#   k_conv_first_s8 1, k_costvol 1, k_cost3d 1 (conv3d_alone).  With --no-e2e the bench creates one context: its eager warm-up pass comes first.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
# 1. every launch with its device time (cold cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 85 -c 425 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_launches.log 2>&1
# 2. full captures (skip the create pass, land on the named layers of the first bench pass)
cap() {   # name, kernel regex, skip, count, bench args
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/${TAG}_$1 $5 > gpurun_out/${TAG}_$1.log 2>&1
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_$1.csv 2>/dev/null
}
cap resblock k_resblock_tc 36 2 "$B"            # refine.2.blocks.0/.1 (full resolution)
cap stream k_conv_stream 59 2 "$B"              # layer2.1.conv_a / conv_b: the shape of 30 of the 55 launches
cap stream3d k_conv_stream 103 1 "$B"           # head.filter.1 (3-D)
cap cost3d k_cost3d 1 1 "$B"                    # conv3d_alone (read-once Conv3d 32 -> 1)
cap refinehead k_refine_head 5 1 "$B"           # full-resolution refinement head (M4/M5 fusion: HBM GB/s)
cap refinehead0 k_refine_head 3 1 "$B"          # stage-0 head: soft-argmin over D + upsample + conv_in
cap costvol k_costvol 1 1 "$B"
# D = 192 (config 4, 4 pairs per pass): cost-volume build and head.filter.0 (its two streaming launches) where the volume no longer fits in L2
B4="python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
cap costvol_d192 k_costvol 1 1 "$B4"
cap filter0_d192 k_conv_stream 101 2 "$B4"
rm -f gpurun_out/${TAG}_*.ncu-rep
# 3. plain runs: per-op CUDA-event table, role counters
python tools/opprof.py --precision tc > gpurun_out/${TAG}_opprof.txt 2>&1
python tools/opprof.py --precision tc --batch 8 > gpurun_out/${TAG}_opprof_b8.txt 2>&1
SNB_TC_PROF=1 python tools/opprof.py --precision tc --reps 1 2>&1 | grep -E "rbprof|csprof|tcprof" > gpurun_out/${TAG}_roleprof.txt
ls -la gpurun_out/ | grep ${TAG}_ | tail -30
