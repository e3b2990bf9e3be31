#!/usr/bin/env python
"""Characterise the accumulate arithmetic of tcgen05.mma kind::f16 (fp32 accumulators in TMEM) on a B200.

Drives tools/ubench/libmma_round_probe.so (mma_round_probe.cu).  Experiments (every row of the 128 x 16 result is one
case; a step is one MMA with K = 16):
  E1  c + ONE product p, p swept in steps of ulp(c)/128, all sign combinations  -> guard bits, truncation mode
  E2  c + 16 EQUAL products                                                   -> per-addend or per-sum truncation
  E3  c + 16 random products of mixed sign and magnitude                      -> model check
  E4  small c, large products (alignment to the largest product)
  E5  chains of 24 accumulating MMAs of random non-negative data              -> bias per MMA in ulp(acc)
A family of hardware models (guard bits g, addend truncation RZ / floor / RN, result rounding, products summed in
groups of G) is evaluated exactly (integer arithmetic) against every observed value; the matching models are printed.
"""
import ctypes as C
import itertools
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
S = 120                                  # exact arithmetic: integers scaled by 2^S


def to_int(x):                           # exact scaled integer of a float
    import math
    m, e = math.frexp(float(x))
    mi = int(m * float(1 << 53))
    return mi << (S + e - 53) if S + e - 53 >= 0 else mi >> (53 - S - e)


def ilog2(v):                            # floor(log2 |v|) of a scaled integer (in real units), None for 0
    return None if v == 0 else abs(v).bit_length() - 1 - S


def trunc_to(v, unit_log2, mode):
    """Round the scaled integer v to a multiple of 2^unit_log2: mode 'rz' (toward zero), 'floor', 'rn' (ties to even)."""
    sh = unit_log2 + S
    if sh <= 0:
        return v
    u = 1 << sh
    if mode == "floor":
        return (v >> sh) << sh
    if mode == "rz":
        return (abs(v) >> sh << sh) * (1 if v >= 0 else -1)
    q, r = divmod(v, u)                  # floor div
    if r * 2 > u or (r * 2 == u and (q & 1)):
        q += 1
    return q * u


def round_f32(v, mode):
    e = ilog2(v)
    if e is None:
        return 0
    return trunc_to(v, e - 23, mode)


def model_step(c, prods, g, tmode, rmode, G):
    acc = c
    for i in range(0, len(prods), G):
        grp = prods[i:i + G]
        es = [ilog2(x) for x in [acc] + grp if x != 0]
        if not es:
            continue
        unit = max(es) - 23 - g
        acc = round_f32(sum(trunc_to(x, unit, tmode) for x in [acc] + grp), rmode)
    return acc


def run(lib, A, B):
    nsteps, _, _ = A.shape
    N = B.shape[1]
    D = np.zeros((128, N), np.float32)
    a16, b16 = np.ascontiguousarray(A, np.float16), np.ascontiguousarray(B, np.float16)
    assert (a16.astype(np.float64) == A).all() and (b16.astype(np.float64) == B).all(), "operands must be exact in fp16"
    r = lib.mma_round_probe(a16.ctypes.data, b16.ctypes.data, nsteps, N, D.ctypes.data)
    if r != 0:
        raise SystemExit(f"probe failed: {r}")
    return D


def collect(out_path):
    """On the GPU box: run every experiment, save operands + results for the offline fit."""
    lib = C.CDLL(os.path.join(HERE, "libmma_round_probe.so"))
    lib.mma_round_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(7)
    cases = []                           # (name, A [steps,128,16], B [steps,16,16])

    # E1: c = +-64 (ulp 2^-17), one product p = ja * 2^-24 = ja/128 ulp; columns: sign of c x sign of b
    ja = np.concatenate([np.arange(-64, 64), ])          # 128 rows
    A = np.zeros((2, 128, 16)); B = np.zeros((2, 16, 16))
    A[0, :, 0] = 8.0
    A[1, :, 0] = ja * 2.0 ** -10
    for n in range(16):
        B[0, n, 0] = 8.0 if (n & 1) == 0 else -8.0
        B[1, n, 0] = (2.0 ** -14) * (1 if (n & 2) == 0 else -1) * (1 + (n >> 2))     # p scaled by 1..4
    cases.append(("E1 single product", A.copy(), B.copy()))
    # E2: 16 equal products, each ja/128/16 ulp ... use b = 2^-14 and a = ja * 2^-10 in all 16 k
    A2 = A.copy(); B2 = B.copy()
    A2[1, :, :] = (ja * 2.0 ** -10)[:, None]
    B2[1, :, :] = B[1, :, 0:1]
    cases.append(("E2 sixteen equal products", A2, B2))
    # E3: random mixed products
    A3 = np.zeros((2, 128, 16)); B3 = np.zeros((2, 16, 16))
    A3[0, :, 0] = 8.0; B3[0, :, 0] = np.where(np.arange(16) % 2 == 0, 8.0, -8.0)
    A3[1] = rng.integers(-2047, 2048, (128, 16)) * 2.0 ** -10
    B3[1] = rng.integers(-15, 16, (16, 16)) * 2.0 ** -14
    cases.append(("E3 random products, large c", A3, B3))
    # E4: small c, products dominate
    A4 = A3.copy(); B4 = B3.copy()
    A4[0, :, 0] = 2.0 ** -10; B4[0, :, 0] = np.where(np.arange(16) % 2 == 0, 2.0 ** -10, -2.0 ** -10) * (1 + np.arange(16) // 2)
    B4[1] = rng.integers(-15, 16, (16, 16)) * 2.0 ** -8
    cases.append(("E4 random products, small c", A4, B4))
    # E3b: products spanning many binades
    A6 = A3.copy(); B6 = B3.copy()
    B6[1] = rng.integers(-15, 16, (16, 16)) * 2.0 ** rng.integers(-14, -2, (16, 16))
    cases.append(("E3b products over 12 binades", A6, B6))

    observed = []
    for name, A_, B_ in cases:
        D = run(lib, A_, B_)
        observed.append((name, A_, B_, D))
    save = {}
    for i, (name, A_, B_, D) in enumerate(observed):
        save[f"name{i}"], save[f"A{i}"], save[f"B{i}"], save[f"D{i}"] = name, A_, B_, D
    ja = np.arange(-64, 64)

    # raw view of E1, column 0 (c = +64, b = +2^-14) and column 1 (c = -64): result - c in units of ulp/128
    name, A_, B_, D = observed[0]
    print("E1: c = +-64, p = ja/128 ulp(c).  (D - c)/ulp for ja = -64..63:")
    for n in (0, 1, 2, 3):
        c = 64.0 if n % 2 == 0 else -64.0
        sgn = 1 if (n & 2) == 0 else -1
        print(f"  col {n}: c = {c:+.0f}, p = {sgn:+d} * ja/128 ulp:", " ".join(f"{(D[m, n] - c) / 2.0 ** -17:+.0f}" for m in range(0, 128, 8)))
    # thresholds: smallest |p| (in ulp/128) at which the result moves, per sign combination
    for n in range(4):
        c = 64.0 if n % 2 == 0 else -64.0
        moved = [(int(ja[m]), (D[m, n] - c) / 2.0 ** -17) for m in range(128)]
        up = [j for j, d in moved if d != 0]
        print(f"  col {n}: result differs from c for ja in {sorted(up)[:3]} ... {sorted(up)[-3:] if up else ''} ({len(up)} of 128)")

    # E5: chains - bias per MMA in units of ulp(acc) for non-negative data (post-ReLU activations x mixed weights, and all-positive)
    for ci, (label, wsign, nsteps) in enumerate((("mixed-sign weights", True, 24), ("positive weights", False, 24), ("mixed-sign weights", True, 6))):
        A5 = rng.integers(0, 2048, (nsteps, 128, 16)) * 2.0 ** -11
        B5 = rng.integers(-1023 if wsign else 0, 1024, (nsteps, 16, 16)) * 2.0 ** -12
        D5 = run(lib, A5, B5)
        save[f"cA{ci}"], save[f"cB{ci}"], save[f"cD{ci}"] = A5, B5, D5
        exact = np.einsum("smk,snk->mn", A5, B5)
        err = D5.astype(np.float64) - exact
        ulp = 2.0 ** (np.floor(np.log2(np.maximum(np.abs(exact), 1e-30))) - 23)
        sel = np.abs(exact) > 1e-3
        print(f"E5 chain of {nsteps} MMAs, {label}: mean error {np.mean(err[sel] / ulp[sel]):+.3f} ulp(final), "
              f"rms {np.sqrt(np.mean((err[sel] / ulp[sel]) ** 2)):.3f} ulp, per MMA {np.mean(err[sel] / ulp[sel]) / nsteps:+.4f} ulp; "
              f"mean |exact| {np.abs(exact[sel]).mean():.3f}")
    np.savez_compressed(out_path, **save)
    print("saved", out_path)


def fit(path):
    """Offline (CPU): evaluate the model family against the saved observations."""
    z = np.load(path)
    observed = []
    i = 0
    while f"A{i}" in z:
        observed.append((str(z[f"name{i}"]), z[f"A{i}"], z[f"B{i}"], z[f"D{i}"]))
        i += 1
    best = []
    for g, tmode, rmode, G in itertools.product(range(0, 10), ("rz", "floor", "rn"), ("rz", "floor", "rn"), (4, 8, 16)):
        bad = tot = 0
        for name, A_, B_, D in observed:
            for m in range(0, 128, 3):
                for n in range(16):
                    c = to_int(A_[0, m, 0]) * to_int(B_[0, n, 0]) >> S
                    prods = [to_int(A_[1, m, k]) * to_int(B_[1, n, k]) >> S for k in range(16)]
                    want = model_step(round_f32(c, "rn"), prods, g, tmode, rmode, G)
                    tot += 1
                    if want != to_int(D[m, n]):
                        bad += 1
        best.append((bad, tot, g, tmode, rmode, G))
    best.sort()
    print("model fit (mismatches / cases, guard bits, addend truncation, result rounding, group size):")
    for b in best[:12]:
        print("  ", b)

    # the best model replayed on the saved chains
    _, _, g, tmode, rmode, G = best[0]
    ci = 0
    while f"cA{ci}" in z:
        A5, B5, D5 = z[f"cA{ci}"], z[f"cB{ci}"], z[f"cD{ci}"].astype(np.float64)
        rows = np.arange(0, 128, 8)
        pred = np.zeros((len(rows), 16))
        for ri, m in enumerate(rows):
            for n in range(16):
                acc = 0
                for s_ in range(A5.shape[0]):
                    prods = [to_int(A5[s_, m, k]) * to_int(B5[s_, n, k]) >> S for k in range(16)]
                    acc = model_step(acc, prods, g, tmode, rmode, G)
                pred[ri, n] = acc / 2.0 ** S
        print(f"chain {ci} ({A5.shape[0]} MMAs): best model reproduces {(pred == D5[rows]).mean() * 100:.1f} % of {pred.size} results exactly")
        ci += 1


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--fit":
        fit(sys.argv[2])
    else:
        collect(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/mma_round_probe.npz")
