"""TEST INFRASTRUCTURE — CPU oracle, never imported by the product path.

Canonical StereoNet architecture reconstructed from the reference's compiled
BPU model (`stereonet_infer/config/hobot_stereonet.hbm`, tensor table at
hbm@0x1b60).  The reference ships the network only as that
binary, so what is restated here is the *topology and shapes* the tensor table
proves, with the free choices (stride placement, dilation, activation) fixed
and documented in DESIGN.md §2.  PARITY UNPINNED for the network: the
reference holds no golden input/output pair for it.

The single source of truth for layer names/shapes; `weights.py` generates the
blob from it, `stereonet_ref.py` runs it in fp32, and the CUDA host code
(`hobot_stereonet_b200/csrc/net.cu`) looks tensors up by the same names.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

# wire-format constants pinned by the reference
OUT_SCALE = 2.60443857769133e-06   # hbm@0x27be8; publisher_member_function.py:29
OUT_NORM = 192.0                   # parser.cpp:86 (16*12); publisher_member_function.py:75
IN_SCALE = 1.0 / 128.0             # preprocess.cpp:1037 ((x-128)/128)

GWC_CH = 256        # layer3 ⊕ layer4 feature (hbm rec [20] 45x80x256)
GWC_GROUPS = 32     # 8 channels per group (SURVEY §2.3 "group-corr volume")
CAT_CH = 16         # lastconv_1 bias valid = 16 (hbm rec [369])
VOL_CH = 2 * CAT_CH + GWC_GROUPS   # 64 (hbm rec [33])
AGG_CH = 32         # _head_filter_* width
REF_CH = 32         # refinement width
REF_DILATIONS = (1, 2, 4, 8, 1, 1)  # StereoNet paper; "astrous" blocks in the hbm names
LAYER_BLOCKS = (3, 16, 3, 3)        # _backbone_layer{1..4}_*
LAYER_CH = (32, 64, 128, 128)


@dataclass(frozen=True)
class Config:
    H: int          # valid (un-padded) image height
    W: int          # valid image width
    K: int          # number of x2 refinement stages == log2(cost-volume stride)
    D: int          # disparity hypotheses at cost-volume resolution

    @property
    def stride(self) -> int:
        return 1 << self.K

    @property
    def Hp(self) -> int:
        s = self.stride
        return (self.H + s - 1) // s * s

    @property
    def Wp(self) -> int:
        s = self.stride
        return (self.W + s - 1) // s * s

    @property
    def h(self) -> int:
        return self.Hp // self.stride

    @property
    def w(self) -> int:
        return self.Wp // self.stride

    @property
    def max_disp(self) -> int:
        return self.stride * self.D


@dataclass(frozen=True)
class ConvSpec:
    name: str
    cout: int
    cin: int
    k: Tuple[int, ...]     # (3,3), (1,1) or (3,3,3)
    stride: int = 1
    dil: int = 1
    relu: bool = True
    gain: float = 1.0      # multiplier on the He-normal std (weights.py)


def layer_strides(K: int) -> Tuple[int, int, int, int]:
    """Backbone stride schedule: firstconv /4 always; layer2 x2 iff K>=3;
    layer3 x2 iff K>=4 (the deployed K=4 model: 180x320 -> 90x160 -> 45x80)."""
    return (1, 2 if K >= 3 else 1, 2 if K >= 4 else 1, 1)


def conv_specs(K: int) -> List[ConvSpec]:
    """Every convolution of the network in execution order (one siamese branch)."""
    out: List[ConvSpec] = []
    # firstconv: hbm weights [41,43,45] 32x(3x3)x3, 32x9x32, 32x9x32, total stride 4
    out.append(ConvSpec("backbone.firstconv.0", 32, 3, (3, 3), stride=2))
    out.append(ConvSpec("backbone.firstconv.1", 32, 32, (3, 3)))
    out.append(ConvSpec("backbone.firstconv.2", 32, 32, (3, 3), stride=2))
    strides = layer_strides(K)
    cin = 32
    for li, (nb, ch, st) in enumerate(zip(LAYER_BLOCKS, LAYER_CH, strides), start=1):
        dil = 2 if li == 4 else 1
        for b in range(nb):
            s = st if b == 0 else 1
            p = f"backbone.layer{li}.{b}"
            out.append(ConvSpec(p + ".conv_a", ch, cin if b == 0 else ch, (3, 3), stride=s, dil=dil))
            # conv_b is the "short_add" conv: + shortcut, then ReLU; damped so 25 blocks stay O(1)
            out.append(ConvSpec(p + ".conv_b", ch, ch, (3, 3), dil=dil, relu=False, gain=0.35))
            if b == 0 and li in (1, 2, 3):
                out.append(ConvSpec(p + ".downsample", ch, cin, (1, 1), stride=s, relu=False, gain=0.7))
        cin = ch
    out.append(ConvSpec("backbone.lastconv.0", 128, GWC_CH, (3, 3)))
    out.append(ConvSpec("backbone.lastconv.1", CAT_CH, 128, (1, 1), relu=False, gain=0.7))
    out.append(ConvSpec("head.filter.0", AGG_CH, VOL_CH, (3, 3, 3)))
    for i in range(1, 5):
        out.append(ConvSpec(f"head.filter.{i}", AGG_CH, AGG_CH, (3, 3, 3)))
    out.append(ConvSpec("head.conv3d_alone", 1, AGG_CH, (3, 3, 3), relu=False, gain=6.0))
    for s in range(K):
        p = f"head.refine.{s}"
        out.append(ConvSpec(p + ".conv_in", REF_CH, 4, (3, 3)))
        for b, d in enumerate(REF_DILATIONS):
            out.append(ConvSpec(f"{p}.blocks.{b}.conv_a", REF_CH, REF_CH, (3, 3), dil=d))
            out.append(ConvSpec(f"{p}.blocks.{b}.conv_b", REF_CH, REF_CH, (3, 3), dil=d, relu=False, gain=0.35))
        out.append(ConvSpec(p + ".conv_out", 1, REF_CH, (3, 3), relu=False, gain=0.02))
    return out


def macs_per_pair(cfg: Config) -> dict:
    """Algorithmic multiply-accumulates per stereo pair, by stage (both siamese branches)."""
    K = cfg.K
    res = {"backbone": 0, "costvol": 0, "agg3d": 0, "refine": 0}
    hh, ww = cfg.Hp, cfg.Wp
    strides = layer_strides(K)
    cur = [hh, ww]

    def conv(spec: ConvSpec, h, w):
        ho, wo = (h + spec.stride - 1) // spec.stride, (w + spec.stride - 1) // spec.stride
        kk = 1
        for t in spec.k:
            kk *= t
        return ho, wo, ho * wo * spec.cout * spec.cin * kk

    h, w = cur
    for spec in conv_specs(K):
        if spec.name.startswith("backbone"):
            if spec.name.endswith(".downsample"):
                # same input as the block's conv_a (already advanced): recompute from output size
                res["backbone"] += 2 * h * w * spec.cout * spec.cin
                continue
            h, w, m = conv(spec, h, w)
            res["backbone"] += 2 * m
        elif spec.name.startswith("head.filter") or spec.name == "head.conv3d_alone":
            res["agg3d"] += cfg.D * cfg.h * cfg.w * spec.cout * spec.cin * 27
        elif spec.name.startswith("head.refine"):
            s = int(spec.name.split(".")[2])
            hs, ws = cfg.h << (s + 1), cfg.w << (s + 1)
            res["refine"] += hs * ws * spec.cout * spec.cin * 9
    res["costvol"] = cfg.D * cfg.h * cfg.w * GWC_CH
    res["total"] = sum(res.values())
    return res
