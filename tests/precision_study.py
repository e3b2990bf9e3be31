"""Not a test: CPU study of which stages tolerate plain fp16 tensor-core operands (1 MMA) and which
need the split hi+lo operands (3 MMAs).  Simulates operand rounding inside the fp32 oracle.
usage: python tests/precision_study.py [H W K D]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import arch, prepost_ref as pp, synth, weights  # noqa: E402
from oracle.stereonet_ref import Oracle  # noqa: E402


def main():
    H, W, K, D = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else (136, 240, 3, 24)
    cfg = arch.Config(H, W, K, D)
    wts = weights.generate(K, seed=1234)
    frame = synth.frame(H, W, cfg.max_disp, seed=1235)
    s8 = pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(frame, H, 2 * W), W, H)
    ref = Oracle(cfg, wts).forward_px(s8)
    h16 = lambda t: t.half().float()
    b16 = lambda t: t.bfloat16().float()

    def only(prefixes, f):
        return lambda t, name: f(t) if any(name.startswith(p) for p in prefixes) else t

    cases = {
        "fp16 everywhere": only([""], h16),
        "fp16 backbone only": only(["backbone"], h16),
        "fp16 backbone.layer2 only": only(["backbone.layer2"], h16),
        "fp16 costvol only": only(["costvol"], h16),
        "fp16 agg3d only": only(["head.filter", "head.conv3d"], h16),
        "fp16 refine only": only(["head.refine"], h16),
        "fp16 refine.(K-1) only": only([f"head.refine.{K - 1}"], h16),
        "bf16 refine only": only(["head.refine"], b16),
        "refine: fp16 weights, exact act": lambda t, n: h16(t) if n.startswith("head.refine") and n.endswith(":w") else t,
        "refine: fp16 act, exact weights": lambda t, n: h16(t) if n.startswith("head.refine") and n.endswith(":a") else t,
        "refine.(K-1): fp16 act, exact weights": lambda t, n: h16(t) if n.startswith(f"head.refine.{K - 1}") and n.endswith(":a") else t,
        "refine.(K-1).blocks: fp16 act, exact weights": lambda t, n: h16(t) if n.startswith(f"head.refine.{K - 1}.blocks") and n.endswith(":a") else t,
        "refine.*.blocks: fp16 act, exact weights": lambda t, n: h16(t) if n.startswith("head.refine") and ".blocks." in n and n.endswith(":a") else t,
        "refine.(K-1).blocks: fp16 weights, exact act": lambda t, n: h16(t) if n.startswith(f"head.refine.{K - 1}.blocks") and n.endswith(":w") else t,
        "agg3d: fp16 weights, exact act": lambda t, n: h16(t) if n.startswith("head.filter") and n.endswith(":w") else t,
        "agg3d: fp16 act, exact weights": lambda t, n: h16(t) if n.startswith("head.filter") and n.endswith(":a") else t,
        "backbone: fp16 weights, exact act": lambda t, n: h16(t) if n.startswith("backbone") and n.endswith(":w") else t,
        "backbone: fp16 act, exact weights": lambda t, n: h16(t) if n.startswith("backbone") and n.endswith(":a") else t,
    }
    for name, rf in cases.items():
        got = Oracle(cfg, wts, round_fn=rf).forward_px(s8)
        e = np.abs(got - ref)
        print(f"{name:32s} mean EPE {e.mean():.3e} px   max {e.max():.3e}")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    main()
