"""CPU: pins oracle/prepost_ref.py against the constants and byte layouts the reference itself
fixes and against a literal loop-level transcription of its algorithm."""
import numpy as np
import pytest

from oracle import arch, prepost_ref as pp, synth


def _literal_yuv420_to_yuv444(buf, w, h):
    """Loop-for-loop reading of preprocess.h:128-155 (small sizes only)."""
    out = np.zeros(3 * w * h, np.uint8)
    src_u, src_v = w * h, w * h + w * h // 4
    out[: w * h] = buf[: w * h]
    des_u, des_v = w * h, 2 * w * h
    for i in range(0, h, 2):
        for j in range(0, w, 2):
            u = buf[src_u + i // 2 * w // 2 + j // 2]
            v = buf[src_v + i // 2 * w // 2 + j // 2]
            for di in (0, 1):
                for dj in (0, 1):
                    out[des_u + (i + di) * w + j + dj] = u
                    out[des_v + (i + di) * w + j + dj] = v
    return out.reshape(3, h, w)


def test_quantize_identity_all_bytes():
    # preprocess.cpp:1037 + :1131-1136 with preprocess.h:236-240 defaults == x - 128 for every byte
    x = np.arange(256, dtype=np.float32)
    q = pp.quantize((x - np.float32(128.0)) / np.float32(128.0))
    assert q.dtype == np.int8
    assert (q.astype(np.int32) == np.arange(256) - 128).all()
    assert ((np.arange(256, dtype=np.uint8) ^ 0x80).view(np.int8) == q).all()


def test_quantize_rounding_and_clamp():
    assert pp.quantize(1.0) == 127 and pp.quantize(-2.0) == -128      # clamp [-128,127]
    assert pp.quantize(0.0039) == 0 and pp.quantize(0.004) == 1       # floor(v/scale + 0.5)
    assert pp.quantize(-0.0039063) == -1 and pp.quantize(-0.00390625) == 0


def test_yuv444_quirk_matches_literal_loops():
    rng = np.random.default_rng(0)
    for (h, w) in [(4, 8), (6, 4), (16, 12)]:
        buf = rng.integers(0, 256, h * w * 3 // 2, dtype=np.uint8)
        assert (pp.yuv420_to_yuv444(buf, w, h) == _literal_yuv420_to_yuv444(buf, w, h)).all()


def test_yuv444_quirk_differs_from_correct_chroma():
    rng = np.random.default_rng(1)
    h, w = 8, 8
    buf = rng.integers(0, 256, h * w * 3 // 2, dtype=np.uint8)
    a, b = pp.yuv420_to_yuv444(buf, w, h), pp.yuv420_to_yuv444(buf, w, h, correct_chroma=True)
    assert (a[0] == b[0]).all() and not (a[1:] == b[1:]).all()
    uv = buf[h * w:].reshape(h // 2, w // 2, 2)
    assert (b[1][::2, ::2] == uv[..., 0]).all() and (b[2][::2, ::2] == uv[..., 1]).all()


def test_split_and_tensor_layout():
    h, w = 8, 16
    frame = synth.frame(h, w, 16, seed=3)
    assert frame.size == h * 3 // 2 * 2 * w
    left, right = pp.split_side_by_side_nv12(frame, h, 2 * w)
    rows = frame.reshape(h * 3 // 2, 2 * w)
    assert (left.reshape(-1, w) == rows[:, :w]).all() and (right.reshape(-1, w) == rows[:, w:]).all()
    t = pp.cvt_nv12_to_tensor(left, right, w, h)
    assert t.shape == (1, 6, h, w) and t.dtype == np.int8
    assert (t == pp.cvt_nv12_to_tensor_fast(left, right, w, h)).all()
    # channels 0-2 left planes, 3-5 right planes (preprocess.cpp:999-1003); Y = y - 128
    assert (t[0, 0].astype(np.int32) + 128 == rows[:h, :w]).all()
    assert (t[0, 3].astype(np.int32) + 128 == rows[:h, w:]).all()


def test_output_scale_and_payload_layout():
    # hbm@0x27be8 scale; x16x12 (parser.cpp:86; publisher_member_function.py:73-75)
    assert arch.OUT_SCALE == 2.60443857769133e-06 and arch.OUT_NORM == 16 * 12
    h, w = 720, 1280
    q = np.full((1, 1, h, w), 200000, np.int32)
    jpeg = b"\xff\xd8fakejpeg\xff\xd9"
    payload = pp.pack_output(q, jpeg)
    assert len(payload) == h * w * 4 + len(jpeg)            # step = len(data), node.cpp:1042
    q2, j2 = pp.unpack_output(payload, h, w)
    assert q2.dtype == np.uint32 and (q2 == 200000).all() and j2 == jpeg
    px = pp.disparity_px(q2)
    assert np.allclose(px, 200000 * 2.60443857769133e-06 * 192)


def test_depth_formula():
    q = np.array([[383960]], np.int32)                     # ~192 px
    px = pp.disparity_px(q)
    z = pp.depth_m(q)
    assert np.allclose(z, 527.1931762695312 * 119.89382172 / px / 1000)
    z32 = pp.parse_tensor_depth_f32(q)
    assert z32.dtype == np.float32 and np.allclose(z32, z, rtol=1e-6)
    with np.errstate(divide="ignore"):
        assert np.isinf(pp.parse_tensor_depth_f32(np.zeros((1, 1), np.int32))).all()   # disp 0 -> inf, as the C++


def test_golden_prepost_fixture():
    g = np.load(__file__.rsplit("/", 1)[0] + "/golden/prepost_16x24.npz")
    left, right = pp.split_side_by_side_nv12(g["frame"], 16, 48)
    assert (pp.cvt_nv12_to_tensor(left, right, 24, 16) == g["s8"]).all()
    assert (pp.cvt_nv12_to_tensor(left, right, 24, 16, correct_chroma=True) == g["s8_correct"]).all()


def test_render_depth_colormap_semantics():
    """parser.cpp:79-118: depth -> convertScaleAbs(alpha) -> JET, through cv2 itself; q = 0 (depth = inf) maps to bin 0."""
    import cv2
    q = np.array([[0, 1, 1000, 383962, 10 ** 6, 2 ** 31 - 1]], np.int32)
    depth, bgr = pp.render_depth_colormap(q, alpha=11.0)
    assert np.isinf(depth[0, 0]) and abs(depth[0, 3] - 0.3292) < 1e-3
    lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, -1), cv2.COLORMAP_JET)[0]
    assert (bgr[0, 0] == lut[0]).all()                    # inf does not fit an int32: bin 0 on x86
    assert (bgr[0, 3] == lut[4]).all()                    # 0.3292 m * 11 = 3.62 -> 4
    assert (bgr[0, 1] == lut[255]).all() and (bgr[0, 2] == lut[255]).all()   # 1.26e5 m and 126 m saturate
    assert (bgr[0, 5] == lut[0]).all()                                        # 5.9e-5 m * 11 rounds to 0


def test_committed_jet_table_matches_cv2():
    import cv2, os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "hobot_stereonet_b200", "csrc", "jet_lut.inc")).read()
    vals = np.array([int(v) for v in re.findall(r"\d+", txt.split("\n", 1)[1])], np.uint8).reshape(256, 3)
    lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, -1), cv2.COLORMAP_JET)[0]
    assert (vals == lut).all()
