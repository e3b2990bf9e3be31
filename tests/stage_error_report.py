"""Not a test: per-stage error of the CUDA path against the CPU oracle at a given size (GPU needed).
usage: python tests/stage_error_report.py H W K D [--f64] [--modes fp32,split,tc] [--json out.json]

Prints max-abs / scale and mean-abs per stage for the fp32 CUDA-core path, split-fp16 storage on CUDA cores, and the
tcgen05 path.  With --f64 the same float model is also evaluated in float64 (exact to ~1e-16: the fp32 weights and
inputs are exactly representable), and every column is reported against BOTH the fp32 oracle ("the reference float
model" of BASELINE.json's metric) and the float64 value, next to the fp32 oracle's own rounding noise
(oracle fp32 vs float64) - the floor no fp32-class implementation can be told apart from."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import arch, prepost_ref as pp, synth, weights  # noqa: E402
from oracle.stereonet_ref import Oracle  # noqa: E402

STAGES = ["firstconv", "layer1", "layer2", "layer3", "layer4", "cat", "volume", "filter0", "filter1", "filter2", "filter3",
          "filter4", "cost", "disp0", "refine0.feat", "disp1", "disp2", "disp3", "disp4"]


def make_s8(cfg, seed=1235):
    H2, W2 = cfg.H + cfg.H % 2, cfg.W + cfg.W % 2
    frame = synth.frame(H2, W2, cfg.max_disp, seed=seed)
    s8 = pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(frame, H2, 2 * W2), W2, H2)
    return np.ascontiguousarray(s8[:, :, :cfg.H, :cfg.W])


def stage_err(got, ref, max_disp, is_disp):
    e = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    sc = max(1.0, float(np.abs(ref).max()))
    r = {"max_over_scale": float(e.max() / sc), "mean": float(e.mean()), "bias": float((got.astype(np.float64) - ref).mean())}
    if is_disp:
        r["mean_px"] = float(e.mean() * max_disp)
        r["max_px"] = float(e.max() * max_disp)
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("shape", nargs="*", type=int, default=[540, 960, 3, 24])
    ap.add_argument("--f64", action="store_true")
    ap.add_argument("--modes", default="fp32,split,tc")
    ap.add_argument("--json", default=None)
    ap.add_argument("--seed", type=int, default=1235)
    args = ap.parse_args()
    from hobot_stereonet_b200 import Model, capi
    H, W, K, D = args.shape
    cfg = arch.Config(H, W, K, D)
    torch.set_num_threads(os.cpu_count())
    s8 = make_s8(cfg, args.seed)
    wts = weights.generate(K, seed=1234)
    refs = {}
    dump = {}
    Oracle(cfg, wts).forward_norm(s8, dump)
    refs["oracle32"] = {k: v.numpy() for k, v in dump.items()}
    if args.f64:
        dump = {}
        Oracle(cfg, wts, dtype=torch.float64).forward_norm(s8, dump)
        refs["oracle64"] = {k: v.numpy() for k, v in dump.items()}
    out = {"shape": {"H": H, "W": W, "K": K, "D": D, "max_disp": cfg.max_disp}, "modes": {}}
    if args.f64:
        print("== fp32 oracle vs float64 oracle (the reference float model's own rounding noise)")
        rep = {}
        for name in STAGES:
            if name not in refs["oracle32"]:
                continue
            rep[name] = stage_err(refs["oracle32"][name], refs["oracle64"][name], cfg.max_disp, name.startswith("disp"))
            r = rep[name]
            extra = f"  (= {r['mean_px']:.3e} px mean, {r['max_px']:.3e} max)" if "mean_px" in r else ""
            print(f"  {name:14s} max/scale {r['max_over_scale']:.3e}  mean {r['mean']:.3e}  bias {r['bias']:+.3e}{extra}")
        out["oracle32_vs_oracle64"] = rep
    blob = weights.make_blob(K, seed=1234)
    modes = {"fp32": (capi.PREC_FP32, 0), "split": (capi.PREC_TC_F16X2, capi.FLAG_NO_TENSOR), "tc": (capi.PREC_TC_F16X2, 0)}
    for label in args.modes.split(","):
        prec, fl = modes[label]
        m = Model(H, W, K, D, weights=blob, precision=prec, flags=fl | capi.FLAG_KEEP_STAGES)
        m.infer(s8)
        out["modes"][label] = {}
        for rname, ref in refs.items():
            print(f"== {label} vs {rname}")
            rep = {}
            for name in STAGES:
                if name not in ref:
                    continue
                got = m.debug_read(name)
                if got.ndim == ref[name].ndim + 1:
                    got = got[:, 0]
                rep[name] = stage_err(got, ref[name], cfg.max_disp, name.startswith("disp"))
                r = rep[name]
                extra = f"  (= {r['mean_px']:.3e} px mean, {r['max_px']:.3e} max)" if "mean_px" in r else ""
                print(f"  {name:14s} max/scale {r['max_over_scale']:.3e}  mean {r['mean']:.3e}  bias {r['bias']:+.3e}{extra}")
            out["modes"][label][rname] = rep
        m.close()
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
