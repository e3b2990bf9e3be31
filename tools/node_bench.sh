#!/bin/bash
# Run under gpurun [--gpus N]: the C++ host node (hobot_stereonet_b200/lib/stereonet_infer: StereonetNode over DnnNode over the C
# ABI) fed with camera frames on 1..N GPUs of the box; prints frames/s per device count.  usage: tools/node_bench.sh <tag> [H W K D]
set -u
TAG=${1:-r02}; H=${2:-720}; W=${3:-1280}; K=${4:-4}; D=${5:-12}
mkdir -p gpurun_out
python - <<PY
import numpy as np, sys
sys.path.insert(0, ".")
from hobot_stereonet_b200 import capi
open("/tmp/model.snb", "wb").write(capi.synthesize_weights($K, 1234))
rng = np.random.default_rng(0)
rng.integers(0, 256, (16, $H * 3 // 2, 2 * $W), dtype=np.uint8).tofile("/tmp/frames.nv12")
PY
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  [ $n -gt $NG ] && break
  DEVS=$(seq -s, 0 $((n-1)))
  for jpeg in off on; do
    ./hobot_stereonet_b200/lib/stereonet_infer --model_file /tmp/model.snb --frames /tmp/frames.nv12 --out /dev/null --model_in_h $H --model_in_w $W \
      --K $K --D $D --devices $DEVS --jpeg $jpeg --repeat $((16 * n)) 2>&1 | grep -E "frames/s|fail|error" | sed "s/^/devices=$DEVS jpeg=$jpeg: /"
  done
done | tee gpurun_out/${TAG}_node_bench.txt
