"""TEST INFRASTRUCTURE — numpy restatement of the reference's host-side pre- and
post-processing around the model call.  Byte/integer work: the bar is bit-exact.

Pinned against the reference's own constants: see
tests/test_oracle_prepost.py.  Each function cites the reference lines it follows.
"""
from __future__ import annotations

import numpy as np

from .arch import OUT_NORM, OUT_SCALE

CAM_F = 527.1931762695312    # parser.cpp:70; publisher_member_function.py:30
CAM_B = 119.89382172         # parser.cpp:71; publisher_member_function.py:31  (mm)


def split_side_by_side_nv12(frame: np.ndarray, h: int, w2: int):
    """stereonet_node.cpp:702-738 — `frame` is one NV12 image of size h x w2 (w2 = 2*model_w):
    h rows of Y then h/2 rows of interleaved UV, each w2 bytes.  Left = first w2/2 bytes of
    every row, right = last w2/2 bytes."""
    rows = np.asarray(frame, dtype=np.uint8).reshape(h * 3 // 2, w2)
    w = w2 // 2
    return np.ascontiguousarray(rows[:, :w]).ravel(), np.ascontiguousarray(rows[:, w:]).ravel()


def yuv420_to_yuv444(buf: np.ndarray, w: int, h: int, correct_chroma: bool = False) -> np.ndarray:
    """preprocess.h:128-155 `Tools::YUV420TOYUV444` -> planar [3,h,w] u8.

    Reference quirk, restated exactly: the routine indexes the chroma block as planar I420
    (srcU = Y + w*h, srcV = srcU + w*h/4, row pitch w/2) although callers hand it NV12
    (stereonet_node.cpp:740), so both output chroma planes hold alternating U,V samples.
    `correct_chroma=True` de-interleaves NV12 properly instead (not what the reference does)."""
    buf = np.asarray(buf, dtype=np.uint8).ravel()
    y = buf[: w * h].reshape(h, w)
    c = buf[w * h: w * h * 3 // 2]
    if correct_chroma:
        uv = c.reshape(h // 2, w // 2, 2)
        u_s, v_s = uv[..., 0], uv[..., 1]
    else:
        u_s = c[: w * h // 4].reshape(h // 2, w // 2)
        v_s = c[w * h // 4:].reshape(h // 2, w // 2)
    up = lambda p: np.repeat(np.repeat(p, 2, axis=0), 2, axis=1)
    return np.stack([y, up(u_s), up(v_s)])


def quantize(value, scale=0.0078125, zero_point=0.5, lo=-128.0, hi=127.0):
    """preprocess.cpp:1131-1136 `PreProcess::Quantize` with the defaults of preprocess.h:236-240
    (float32 arithmetic, floor rounding)."""
    v = np.floor(np.float32(value) / np.float32(scale) + np.float32(zero_point))
    return np.clip(v, np.float32(lo), np.float32(hi)).astype(np.int8)


def cvt_nv12_to_tensor(left: np.ndarray, right: np.ndarray, w: int, h: int,
                       correct_chroma: bool = False) -> np.ndarray:
    """preprocess.cpp:913-1059 `PreProcess::CvtNV12Data2Tensors` -> s8 NCHW [1,6,h,w].
    L planes then R planes (:999-1003); per byte Quantize((x-128)/128) (:1032-1040)."""
    planes = np.concatenate([yuv420_to_yuv444(left, w, h, correct_chroma),
                             yuv420_to_yuv444(right, w, h, correct_chroma)])
    x = (planes.astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
    return quantize(x)[None]


def cvt_nv12_to_tensor_fast(left, right, w, h, correct_chroma=False) -> np.ndarray:
    """Same result as `cvt_nv12_to_tensor` via the identity Quantize((x-128)/128) == x-128
    (SURVEY.md §4 item 1): byte = u8 ^ 0x80 reinterpreted as s8."""
    planes = np.concatenate([yuv420_to_yuv444(left, w, h, correct_chroma),
                             yuv420_to_yuv444(right, w, h, correct_chroma)])
    return (planes ^ np.uint8(0x80)).view(np.int8)[None]


def pack_output(infer_s32: np.ndarray, jpeg: bytes) -> bytes:
    """stereonet_node.cpp:1033-1049 — payload = raw int32 LE model output, then the JPEG of the
    left view; msg.step = len(payload), msg.encoding = "jpeg"."""
    return np.ascontiguousarray(infer_s32, dtype="<i4").tobytes() + bytes(jpeg)


def unpack_output(payload: bytes, h: int, w: int):
    """publisher_member_function.py:52-66 — consumer-side slicing of the payload."""
    n = w * h * 4
    q = np.frombuffer(payload[:n], dtype=np.uint32).reshape(1, 1, h, w)
    return q, payload[n:]


def disparity_px(q: np.ndarray) -> np.ndarray:
    """publisher_member_function.py:73-75: image_pre = q * scale * 16 * 12 (float64 numpy)."""
    return q * OUT_SCALE * 16 * 12


def depth_m(q: np.ndarray) -> np.ndarray:
    """publisher_member_function.py:81 / parser.cpp:84-86: Z = f*B/disp_px/1000 (metres)."""
    with np.errstate(divide="ignore"):
        return CAM_F * CAM_B / disparity_px(q) / 1000


def parse_tensor_depth_f32(q: np.ndarray, scale: float = OUT_SCALE) -> np.ndarray:
    """parser.cpp:79-87 `ParseTensor` in its own arithmetic: float dis = (float)q * scale;
    result = f*B/(dis*16.0*12.0)/1000.0 evaluated in double then stored as float."""
    dis = q.astype(np.float32) * np.float32(scale)
    fb = np.float64(np.float32(CAM_F) * np.float32(CAM_B))   # float*float stays float (parser.cpp:86)
    with np.errstate(divide="ignore"):
        return (fb / (dis.astype(np.float64) * 16.0 * 12.0) / 1000.0).astype(np.float32)


def render_depth_colormap(q: np.ndarray, alpha: float = 11.0, scale: float = OUT_SCALE):
    """parser.cpp:79-118 (alpha = 11) / publisher_member_function.py:81-82 (alpha = 9): depth in metres, then
    cv::convertScaleAbs(depth, alpha) and cv::applyColorMap(COLORMAP_JET), through the very cv2 functions the
    reference calls.  q [..., H, W] int32 -> (depth float32 [..., H, W], bgr uint8 [..., H, W, 3])."""
    import cv2
    depth = parse_tensor_depth_f32(q, scale)
    flat = depth.reshape(-1, depth.shape[-1])
    u8 = cv2.convertScaleAbs(flat, alpha=alpha)
    bgr = cv2.applyColorMap(u8, cv2.COLORMAP_JET)
    return depth, bgr.reshape(depth.shape + (3,))
