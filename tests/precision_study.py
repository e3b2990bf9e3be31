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


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "--two-mma"):
    torch.set_num_threads(os.cpu_count())
    main()


# ---- what a 2-MMA product would cost in accuracy (the fp8 correction idea, studied, not built) ----------------
# 3-MMA scheme of the kernels: x*w ~ hi*hi + hi*lo + lo*hi (fp16 MMAs, error 2^-22).  The 2-MMA idea keeps hi*hi in fp16 and
# computes BOTH corrections with ONE kind::f8f6f4 MMA of K = 32: [e4m3(x_hi) | e4m3(x_lo * 2^11)] x [e4m3(w_lo * 2^11) ; e4m3(w_hi)],
# i.e. each correction term carries two 3-bit-mantissa roundings: relative error of the product ~ 2^-11 * 2^-3.5.
def two_mma_study(argv):
    import torch.nn.functional as F
    H, W, K, D = [int(v) for v in argv[:4]] if len(argv) >= 4 else (136, 240, 3, 24)
    cfg = arch.Config(H, W, K, D)
    wts = weights.generate(K, seed=1234)
    frame = synth.frame(H, W, cfg.max_disp, seed=1235)
    s8 = pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(frame, H, 2 * W), W, H)
    ref = Oracle(cfg, wts).forward_px(s8)

    def q8(t):                      # e4m3 with a per-tensor power-of-two scale (max -> [128, 256)), as a kernel would store it
        m = float(t.abs().max())
        if m == 0:
            return t
        s = 2.0 ** (7 - np.floor(np.log2(m)))
        return (t * s).clamp(-448, 448).to(torch.float8_e4m3fn).float() / s

    class Two(Oracle):
        def __init__(self, *a, where=(), **kw):
            super().__init__(*a, **kw)
            self.where = where

        def conv(self, x, name, stride=1, dil=1, relu=True, add=None):
            if not any(name.startswith(p) for p in self.where):
                return super().conv(x, name, stride, dil, relu, add)
            w, b = self.w[name + ".weight"], self.w[name + ".bias"]
            pad = dil * (w.shape[-1] // 2)
            f = (lambda a, ww, bb: F.conv3d(a, ww, bb, stride=stride, padding=pad)) if w.dim() == 5 else \
                (lambda a, ww, bb: F.conv2d(a, ww, bb, stride=stride, padding=pad, dilation=dil))
            xh, wh = x.half().float(), w.half().float()
            xl, wl = (x - xh), (w - wh)
            y = f(xh, wh, b) + f(q8(xh), q8(wl), None) + f(q8(xl), q8(wh), None)
            if add is not None:
                y = y + add
            return F.relu(y) if relu else y

    for label, where in (("2-MMA products in the last refinement stage only", (f"head.refine.{K - 1}",)),
                         ("2-MMA products in all refinement stages", ("head.refine",)),
                         ("2-MMA products in the 3-D aggregation", ("head.filter", "head.conv3d")),
                         ("2-MMA products in the backbone", ("backbone",)),
                         ("2-MMA products everywhere", ("",))):
        got = Two(cfg, wts, where=where).forward_px(s8)
        e = np.abs(got - ref)
        print(f"{label:52s} mean EPE {e.mean():.3e} px   max {e.max():.3e}   (max_disp {cfg.max_disp})")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "--two-mma":
    torch.set_num_threads(os.cpu_count())
    two_mma_study(sys.argv[2:])
