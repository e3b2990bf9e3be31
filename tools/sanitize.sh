#!/bin/bash
# Run under gpurun (one GPU): compute-sanitizer over the CUDA path.  usage: tools/sanitize.sh <tag>
#   racecheck + synccheck   small shapes, every tcgen05 kernel family (their correctness is an mbarrier protocol)
#   memcheck                the whole -m gpu test suite
# Logs land in gpurun_out/<tag>_{racecheck,synccheck,memcheck}.log; copy the summaries into profiles/.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
for tool in racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/${TAG}_${tool}.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/${TAG}_${tool}.log
  tail -4 gpurun_out/${TAG}_${tool}.log
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/${TAG}_memcheck.log
tail -6 gpurun_out/${TAG}_memcheck.log
