"""Not a test: per-stage error of the CUDA path against the fp32 oracle at a given size (GPU needed).
usage: python tests/stage_error_report.py [H W K D] — prints max-abs / scale and mean-abs per stage for
fp32, split-fp16 storage on CUDA cores, and the tcgen05 path."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import arch, prepost_ref as pp, synth, weights  # noqa: E402
from oracle.stereonet_ref import Oracle  # noqa: E402

STAGES = ["firstconv", "layer1", "layer2", "layer3", "layer4", "cat", "volume", "filter0", "filter4", "cost", "disp0",
          "refine0.feat", "disp1", "disp2", "disp3", "disp4"]


def main():
    from hobot_stereonet_b200 import Model, capi
    H, W, K, D = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else (540, 960, 3, 24)
    cfg = arch.Config(H, W, K, D)
    torch.set_num_threads(os.cpu_count())
    frame = synth.frame(H, W, cfg.max_disp, seed=1235)
    s8 = pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(frame, H, 2 * W), W, H)
    dump = {}
    Oracle(cfg, weights.generate(K, seed=1234)).forward_norm(s8, dump)
    blob = weights.make_blob(K, seed=1234)
    modes = {"fp32": (capi.PREC_FP32, 0), "split-storage, CUDA cores": (capi.PREC_TC_F16X2, capi.FLAG_NO_TENSOR),
             "tcgen05": (capi.PREC_TC_F16X2, 0)}
    for label, (prec, fl) in modes.items():
        m = Model(H, W, K, D, weights=blob, precision=prec, flags=fl | capi.FLAG_KEEP_STAGES)
        m.infer(s8)
        print(f"== {label}")
        for name in STAGES:
            if name not in dump:
                continue
            ref = dump[name].numpy()
            got = m.debug_read(name)
            if got.ndim == ref.ndim + 1:
                got = got[:, 0]
            e = np.abs(got - ref)
            sc = max(1.0, float(np.abs(ref).max()))
            extra = f"  (= {e.mean() * cfg.max_disp:.3e} px mean)" if name.startswith("disp") else ""
            print(f"  {name:14s} max/scale {e.max() / sc:.3e}  mean {e.mean():.3e}  bias {float((got - ref).mean()):+.3e}{extra}")
        m.close()


if __name__ == "__main__":
    main()
