"""TEST INFRASTRUCTURE — fp32 PyTorch-CPU oracle of the network behind
`DnnNode::Run()` (stereonet_node.cpp:812).  PARITY UNPINNED: the reference
runs this span as a compiled BPU binary and holds no golden tensors for it
(SURVEY.md §8c); this file restates the topology recovered from the .hbm tensor
table with the free choices documented in DESIGN.md §2.  It is
the "reference float model" of BASELINE.json's metric.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.

Stage order:
  s8[B,6,H,W] -> x/128 -> pad -> siamese backbone -> gwc(256)+cat(16) features
  -> cost volume [64,D,h,w] -> 6 Conv3d -> softmax_D -> soft-argmin (normalised)
  -> K x {x2 bilinear, cat left image, 14 convs, +residual, ReLU} -> s32
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .arch import (CAT_CH, GWC_GROUPS, IN_SCALE, LAYER_BLOCKS, OUT_NORM, OUT_SCALE,
                   REF_DILATIONS, Config, layer_strides)

Tensors = Dict[str, torch.Tensor]


def quant_multiplier(cfg: Config) -> np.float32:
    """q = rint(dn * qmul): dn is disparity normalised by max_disp; decoded by the untouched
    render tool as px = q * OUT_SCALE * 192 (publisher_member_function.py:73-75)."""
    return np.float32(cfg.max_disp / (OUT_NORM * OUT_SCALE))


class Oracle:
    def __init__(self, cfg: Config, weights: Dict[str, np.ndarray],
                 round_fn: Optional[Callable[[torch.Tensor, str], torch.Tensor]] = None,
                 dtype: torch.dtype = torch.float32):
        self.cfg = cfg
        # dtype = torch.float64 gives the exact-arithmetic value of the same float model (the fp32 weights and inputs are
        # exactly representable): the yardstick for the rounding noise of the fp32 oracle itself (tests/test_gpu_d192.py)
        self.dtype = dtype
        self.w: Tensors = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in weights.items()}
        # optional operand rounding for precision studies: round_fn(tensor, tag) -> tensor with
        # tag = "<layer name>:a" (activation operand) / "<layer name>:w" (weight operand) / "costvol",
        # e.g. lambda t, tag: t.half().float() if tag.startswith("head.refine") else t
        self.rf = round_fn or (lambda t, name: t)

    # -- building blocks -------------------------------------------------------------------
    def conv(self, x, name, stride=1, dil=1, relu=True, add=None):
        w, b = self.w[name + ".weight"], self.w[name + ".bias"]
        pad = dil * (w.shape[-1] // 2)
        if w.dim() == 5:
            y = F.conv3d(self.rf(x, name + ":a"), self.rf(w, name + ":w"), b, stride=stride, padding=pad)
        else:
            y = F.conv2d(self.rf(x, name + ":a"), self.rf(w, name + ":w"), b, stride=stride, padding=pad, dilation=dil)
        if add is not None:
            y = y + add
        return F.relu(y) if relu else y

    def block(self, x, p, stride, dil, has_ds):
        y = self.conv(x, p + ".conv_a", stride=stride, dil=dil)
        sc = self.conv(x, p + ".downsample", stride=stride, relu=False) if has_ds else x
        return self.conv(y, p + ".conv_b", dil=dil, relu=True, add=sc)

    def backbone(self, img, dump=None):
        """img [N,3,Hp,Wp] -> (gwc feature [N,256,h,w], concat feature [N,16,h,w])."""
        x = self.conv(img, "backbone.firstconv.0", stride=2)
        x = self.conv(x, "backbone.firstconv.1")
        x = self.conv(x, "backbone.firstconv.2", stride=2)
        if dump is not None:
            dump["firstconv"] = x
        strides = layer_strides(self.cfg.K)
        feats = []
        for li in range(1, 5):
            for b in range(LAYER_BLOCKS[li - 1]):
                x = self.block(x, f"backbone.layer{li}.{b}",
                               stride=strides[li - 1] if b == 0 else 1,
                               dil=2 if li == 4 else 1,
                               has_ds=(b == 0 and li <= 3))
            if dump is not None:
                dump[f"layer{li}"] = x
            feats.append(x)
        gwc = torch.cat([feats[2], feats[3]], dim=1)
        y = self.conv(gwc, "backbone.lastconv.0")
        cat = self.conv(y, "backbone.lastconv.1", relu=False)
        return gwc, cat

    def cost_volume(self, gl, gr, cl, cr):
        """[B,64,D,h,w]: ch 0-15 left concat feature, 16-31 right concat feature shifted by d,
        32-63 group-wise correlation (mean over 8-channel groups); zero where x < d
        (hbm `_head_hz_cat_{1..14}`, `_head_hz_mul*`, `_head_hz_mean*`, `_head_{c,gc}_pad_*`)."""
        B, C, h, w = gl.shape
        D = self.cfg.D
        vol = gl.new_zeros(B, 2 * CAT_CH + GWC_GROUPS, D, h, w)
        for d in range(min(D, w)):
            vol[:, :CAT_CH, d, :, d:] = cl[:, :, :, d:]
            vol[:, CAT_CH:2 * CAT_CH, d, :, d:] = cr[:, :, :, :w - d]
            prod = self.rf(gl[:, :, :, d:], "costvol") * self.rf(gr[:, :, :, :w - d], "costvol")
            vol[:, 2 * CAT_CH:, d, :, d:] = prod.view(B, GWC_GROUPS, C // GWC_GROUPS, h, w - d).mean(2)
        return vol

    def aggregate(self, vol, dump=None):
        x = vol
        for i in range(5):
            x = self.conv(x, f"head.filter.{i}")
            if dump is not None:
                dump[f"filter{i}"] = x
        return self.conv(x, "head.conv3d_alone", relu=False)[:, 0]      # [B,D,h,w]

    def soft_argmin(self, cost):
        p = torch.softmax(cost, dim=1)
        d = torch.arange(self.cfg.D, dtype=cost.dtype, device=cost.device).view(1, -1, 1, 1) / self.cfg.D
        return (p * d).sum(1, keepdim=True)                             # [B,1,h,w] in [0,1)

    def refine(self, disp, left, s, dump=None):
        p = f"head.refine.{s}"
        hs, ws = disp.shape[-2] * 2, disp.shape[-1] * 2
        up = F.interpolate(disp, size=(hs, ws), mode="bilinear", align_corners=False)
        img = left if (hs, ws) == tuple(left.shape[-2:]) else \
            F.interpolate(left, size=(hs, ws), mode="bilinear", align_corners=False)
        x = self.conv(torch.cat([up, img], dim=1), p + ".conv_in")
        for b, dil in enumerate(REF_DILATIONS):
            y = self.conv(x, f"{p}.blocks.{b}.conv_a", dil=dil)
            x = self.conv(y, f"{p}.blocks.{b}.conv_b", dil=dil, relu=True, add=x)
        if dump is not None:
            dump[f"refine{s}.feat"] = x
        return self.conv(x, p + ".conv_out", relu=True, add=up)

    # -- whole network ---------------------------------------------------------------------
    @torch.no_grad()
    def forward_norm(self, s8: np.ndarray, dump: Optional[dict] = None) -> torch.Tensor:
        """s8 [B,6,H,W] int8 -> normalised disparity [B,Hp,Wp] float32 (disp_px / max_disp)."""
        cfg = self.cfg
        assert s8.dtype == np.int8 and s8.shape[1:] == (6, cfg.H, cfg.W), s8.shape
        return self.forward_float((torch.from_numpy(s8.astype(np.float32)) * IN_SCALE).to(self.dtype), dump)

    @torch.no_grad()
    def forward_float(self, x: torch.Tensor, dump: Optional[dict] = None) -> torch.Tensor:
        """x [B,6,H,W] float32 (= s8/128) -> normalised disparity [B,Hp,Wp]."""
        cfg = self.cfg
        x = F.pad(x, (0, cfg.Wp - cfg.W, 0, cfg.Hp - cfg.H))
        B = x.shape[0]
        left, right = x[:, :3], x[:, 3:]
        gwc, cat = self.backbone(torch.cat([left, right], 0), dump)
        if dump is not None:
            dump["gwc"], dump["cat"] = gwc, cat
        vol = self.cost_volume(gwc[:B], gwc[B:], cat[:B], cat[B:])
        if dump is not None:
            dump["volume"] = vol
        cost = self.aggregate(vol, dump)
        disp = self.soft_argmin(cost)
        if dump is not None:
            dump["cost"], dump["disp0"] = cost, disp[:, 0]
        for s in range(cfg.K):
            disp = self.refine(disp, left, s, dump)
            if dump is not None:
                dump[f"disp{s + 1}"] = disp[:, 0]
        return disp[:, 0]

    def forward_px(self, s8: np.ndarray) -> np.ndarray:
        """Left-view disparity in pixels, cropped to the valid H x W."""
        dn = self.forward_norm(s8).numpy()
        return dn[:, :self.cfg.H, :self.cfg.W] * dn.dtype.type(self.cfg.max_disp)

    def forward_s32(self, s8: np.ndarray) -> np.ndarray:
        """The model output tensor as the reference reads it (stereonet_node.cpp:1033):
        int32 NCHW [B,1,H,W], value * OUT_SCALE * 192 = disparity in pixels."""
        dn = self.forward_norm(s8).numpy()[:, :self.cfg.H, :self.cfg.W].astype(np.float32)
        q = np.rint(dn * quant_multiplier(self.cfg)).astype(np.int32)
        return q[:, None]
