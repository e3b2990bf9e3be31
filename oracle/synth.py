"""TEST INFRASTRUCTURE — seeded synthetic SceneFlow-shaped stereo pairs.

left  = 4 octaves of band-limited uniform noise, 3 channels, u8 (BGR)
disp  = clip(planar ramp + 3 Gaussian blobs, 0, 0.9*max_disp)
right = backward warp of left by disp (bilinear): right(x) = left(x + disp)  [left-view disparity]
frame = side-by-side NV12 (left | right), the camera message layout the node consumes
        (stereonet_node.cpp:682-690: height == model_h, width == 2*model_w, encoding "nv12").
BGR->NV12 follows `Tools::BGRToNv12` (preprocess.h:56-96): cv2 BGR2YUV_I420, then interleave U,V.
"""
from __future__ import annotations

import cv2
import numpy as np


def _octave_noise(rng, h, w, octaves=4):
    img = np.zeros((h, w, 3), np.float32)
    amp = 1.0
    for o in range(octaves):
        gh, gw = max(2, h >> (octaves + 1 - o)), max(2, w >> (octaves + 1 - o))
        g = rng.uniform(-1, 1, (gh, gw, 3)).astype(np.float32)
        img += amp * cv2.resize(g, (w, h), interpolation=cv2.INTER_CUBIC)
        amp *= 0.6
    img -= img.min()
    img /= max(img.max(), 1e-6)
    return (img * 255).astype(np.uint8)


def disparity_field(rng, h, w, max_disp):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    d = 0.15 * max_disp + 0.25 * max_disp * (yy / h) + 0.1 * max_disp * (xx / w)
    for _ in range(3):
        cy, cx = rng.uniform(0.2, 0.8) * h, rng.uniform(0.2, 0.8) * w
        s = rng.uniform(0.08, 0.2) * min(h, w)
        d += rng.uniform(0.1, 0.3) * max_disp * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
    return np.clip(d, 0, 0.9 * max_disp).astype(np.float32)


def bgr_to_nv12(bgr: np.ndarray) -> np.ndarray:
    h, w = bgr.shape[:2]
    i420 = cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV_I420).ravel()
    y = i420[: h * w]
    u = i420[h * w: h * w * 5 // 4]
    v = i420[h * w * 5 // 4:]
    uv = np.stack([u, v], axis=1).ravel()
    return np.concatenate([y, uv])


def stereo_pair(h: int, w: int, max_disp: int, seed: int):
    """-> (left_bgr u8 [h,w,3], right_bgr u8 [h,w,3], disp_gt f32 [h,w])."""
    rng = np.random.default_rng(seed)
    left = _octave_noise(rng, h, w)
    disp = disparity_field(rng, h, w, max_disp)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    # a point at x in the left view appears at x - d in the right view; sample by backward warp
    right = cv2.remap(left, xx + disp, yy, interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
    return left, right, disp


def side_by_side_nv12(left_bgr: np.ndarray, right_bgr: np.ndarray) -> np.ndarray:
    """One NV12 frame of width 2w holding left | right — the `HbmMsg1080P.data` payload."""
    h, w = left_bgr.shape[:2]
    ln = bgr_to_nv12(left_bgr).reshape(h * 3 // 2, w)
    rn = bgr_to_nv12(right_bgr).reshape(h * 3 // 2, w)
    return np.concatenate([ln, rn], axis=1).ravel()


def frame(h: int, w: int, max_disp: int, seed: int) -> np.ndarray:
    l, r, _ = stereo_pair(h, w, max_disp, seed)
    return side_by_side_nv12(l, r)
