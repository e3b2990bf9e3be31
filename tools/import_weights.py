"""PyTorch checkpoint -> SNB2WGT1 `model_file`.

  python tools/import_weights.py --checkpoint stereonet_float.pth --K 4 --out hobot_stereonet.snb [--map names.json]

Folds every BatchNorm into its convolution (hobot_stereonet_b200/weights_io.py).  `--map` is a JSON object
{layer: {"conv": "<checkpoint prefix>", "bn": "<checkpoint prefix>"}} for checkpoints whose module names differ from this
repo's layer names (`python tools/import_weights.py --list --K 4` prints them with their shapes)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hobot_stereonet_b200 import weights_io  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint")
    ap.add_argument("--K", type=int, default=4)
    ap.add_argument("--out")
    ap.add_argument("--map")
    ap.add_argument("--eps", type=float, default=1e-5)
    ap.add_argument("--list", action="store_true")
    a = ap.parse_args()
    if a.list:
        for name, shape in weights_io.expected_layers(a.K):
            print(name, "x".join(map(str, shape)))
        return 0
    if not a.checkpoint or not a.out:
        ap.error("--checkpoint and --out are required")
    import torch
    ck = torch.load(a.checkpoint, map_location="cpu", weights_only=True)
    for key in ("state_dict", "model", "state_dict_ema"):
        if isinstance(ck, dict) and key in ck and isinstance(ck[key], dict):
            ck = ck[key]
    sd = {k[len("module."):] if k.startswith("module.") else k: v.detach().cpu().numpy() for k, v in ck.items() if hasattr(v, "detach")}
    name_map = json.load(open(a.map)) if a.map else None
    tensors, report = weights_io.import_state_dict(sd, a.K, name_map, a.eps)
    with open(a.out, "wb") as f:
        f.write(weights_io.write_blob(tensors, a.K))
    n = sum(int(np.prod(t.shape)) for t in tensors.values())
    print(f"{a.out}: {len(tensors) // 2} convolutions, {n} parameters; BatchNorm folded into {len(report['folded'])}, "
          f"{len(report['plain'])} without; {len(report['unused'])} checkpoint tensors unused")
    return 0


if __name__ == "__main__":
    sys.exit(main())
