#!/usr/bin/env python
"""Per-op device times of one pass (snb_profile_pass, CUDA events around every launch).
usage: python tools/opprof.py [--precision tc|fp32] [--H 540 --W 960 --K 3 --D 24 --batch 1] [--reps 5]
Writes a table to stdout; used to pick the kernel to optimise next (profiles/*.md cite it)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="tc")
    ap.add_argument("--H", type=int, default=540)
    ap.add_argument("--W", type=int, default=960)
    ap.add_argument("--K", type=int, default=3)
    ap.add_argument("--D", type=int, default=24)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from hobot_stereonet_b200 import Model, capi
    prec = {"tc": capi.PREC_TC_F16X2, "fp32": capi.PREC_FP32}.get(a.precision)
    if prec is None:
        prec = int(a.precision)
    m = Model(a.H, a.W, a.K, a.D, max_batch=a.batch, precision=prec, weights=capi.synthesize_weights(a.K, 1234))
    acc = {}
    order = []
    for _ in range(a.reps):
        for i, (name, ms, fl, by) in enumerate(m.profile_pass(a.batch)):
            key = (i, name)
            if key not in acc:
                acc[key] = [0.0, fl, by]
                order.append(key)
            acc[key][0] += ms / a.reps
    tot = sum(v[0] for v in acc.values())
    print(f"# {a.H}x{a.W} K={a.K} D={a.D} batch={a.batch} precision={a.precision}: {tot:.3f} ms/pass (event-per-launch, no graph)")
    print(f"{'op':44s} {'ms':>8s} {'%':>6s} {'TFLOP/s':>9s} {'GB/s':>8s}")
    for key in order:
        ms, fl, by = acc[key]
        print(f"{key[1]:44s} {ms:8.4f} {100 * ms / tot:6.2f} {fl / ms / 1e9 if ms > 0 else 0:9.2f} {by / ms / 1e6 if ms > 0 else 0:8.1f}")
    m.close()


if __name__ == "__main__":
    main()
