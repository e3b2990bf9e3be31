#!/usr/bin/env python
"""Small-shape workload for compute-sanitizer (racecheck / synccheck / initcheck): every tcgen05 kernel family of the
tensor-core path once, through the C ABI, checked against the CPU oracle so a tool-induced failure is visible.
  compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import arch, prepost_ref as pp, synth, weights  # noqa: E402
from oracle.stereonet_ref import Oracle  # noqa: E402


def main():
    from hobot_stereonet_b200 import Model, capi
    for (H, W, K, D), flags in (((48, 160, 3, 6), 0), ((32, 64, 2, 5), 0), ((48, 160, 3, 6), capi.FLAG_NO_HEADFUSE | capi.FLAG_NO_GRAPH)):
        cfg = arch.Config(H, W, K, D)
        frame = synth.frame(H, W, cfg.max_disp, seed=3)
        s8 = pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(frame, H, 2 * W), W, H)
        m = Model(H, W, K, D, weights=weights.make_blob(K, seed=1234), precision=capi.PREC_TC_F16X2, flags=flags)
        q = m.infer(s8)
        q2 = m.infer_nv12(frame.reshape(1, H * 3 // 2, 2 * W))
        m.close()
        ref = Oracle(cfg, weights.generate(K, seed=1234)).forward_px(s8)
        err = np.abs(q[:, 0].astype(np.float64) * arch.OUT_SCALE * arch.OUT_NORM - ref)
        assert (q == q2).all() and err.mean() < 1e-3, (H, W, K, D, flags, err.mean())
        print(f"{H}x{W} K={K} D={D} flags={flags}: mean EPE {err.mean():.2e} px")


if __name__ == "__main__":
    main()
