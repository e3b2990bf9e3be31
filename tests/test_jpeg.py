"""The JPEG half of the published payload (stereonet_node.cpp:749-803, 1033-1049) and its consumer.

The reference builds it with cv::cvtColor(CV_YUV2BGR_NV12) + cv::imencode(".jpg"); the library restates both
(csrc/jpeg_host.cpp).  Checked here on the CPU: the colour conversion is bit-exact against cv2, the JPEG decodes with
cv2.imdecode (what the untouched render tool calls, publisher_member_function.py:93-95) to the same picture cv2's own
encoder gives within JPEG's loss, and a full 1280x720 payload goes through the render tool's slicing and depth
arithmetic (publisher_member_function.py:52-98) restated line by line."""
import cv2
import numpy as np
import pytest

from oracle import arch, prepost_ref as pp, synth


def _psnr(a, b):
    return 10 * np.log10(255.0 ** 2 / max(1e-12, float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))))


@pytest.mark.parametrize("h,w", [(16, 16), (38, 70), (540, 960), (720, 1280)])
def test_nv12_to_bgr_bit_exact_vs_cv2(built_lib, h, w):
    from hobot_stereonet_b200 import capi
    rng = np.random.default_rng(h * w)
    for nv in (synth.bgr_to_nv12(synth.stereo_pair(h, w, 64, seed=3)[0]), rng.integers(0, 256, h * 3 // 2 * w, dtype=np.uint8)):
        ref = cv2.cvtColor(nv.reshape(h * 3 // 2, w), cv2.COLOR_YUV2BGR_NV12)       # stereonet_node.cpp:775-777
        assert (capi.nv12_to_bgr(nv, w, h) == ref).all()


@pytest.mark.parametrize("h,w", [(2, 2), (16, 16), (38, 70), (540, 960), (720, 1280)])
def test_jpeg_decodes_like_cv2s_own(built_lib, h, w):
    from hobot_stereonet_b200 import capi
    nv = synth.bgr_to_nv12(synth.stereo_pair(h, w, 64, seed=5)[0])
    bgr = cv2.cvtColor(nv.reshape(h * 3 // 2, w), cv2.COLOR_YUV2BGR_NV12)
    jpg = capi.jpeg_encode_nv12(nv, w, h)
    assert jpg[:2] == b"\xff\xd8" and jpg[-2:] == b"\xff\xd9" and jpg[6:10] == b"JFIF"
    im = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_ANYCOLOR)          # publisher_member_function.py:95
    assert im is not None and im.shape == (h, w, 3)
    ok, ref_jpg = cv2.imencode(".jpg", bgr)                                       # stereonet_node.cpp:780-782, default params
    ref = cv2.imdecode(ref_jpg, cv2.IMREAD_ANYCOLOR)
    if h * w >= 256:
        assert _psnr(im, bgr) >= _psnr(ref, bgr) - 0.3                             # as faithful to the source as libjpeg at q95
        assert _psnr(im, ref) >= 45.0                                              # and the two decodes are the same picture
        assert 0.8 < len(jpg) / len(ref_jpg) < 1.25
    # a smaller buffer is never overrun and the needed size is still reported
    buf = np.full(100, 0xAB, np.uint8)
    n = capi.lib().snb_jpeg_encode_nv12(nv.ctypes.data, w, h, 0, buf.ctypes.data, 50)
    assert n == len(jpg) and (buf[50:] == 0xAB).all()


def test_jpeg_rejects_bad_arguments(built_lib):
    from hobot_stereonet_b200 import capi
    nv = np.zeros(3 * 5 * 3 // 2 + 8, np.uint8)
    assert capi.lib().snb_jpeg_encode_nv12(nv.ctypes.data, 3, 4, 0, None, 0) == capi.SNB_ERR_INVALID      # odd width
    assert capi.lib().snb_jpeg_encode_nv12(None, 4, 4, 0, None, 0) == capi.SNB_ERR_INVALID


def test_payload_through_render_tool_arithmetic(built_lib):
    """publisher_member_function.py:52-98 on a payload built the way the node builds it (snb_post_pack of the s32 tensor
    and the library's JPEG), at the only size the render tool supports (1280 x 720, hard-coded at :72)."""
    from hobot_stereonet_b200 import capi
    h, w = 720, 1280
    left, _, disp = synth.stereo_pair(h, w, 192, seed=11)
    q = np.rint(disp / (arch.OUT_SCALE * arch.OUT_NORM)).astype(np.int32)[None, None]      # what the model would output
    nv = synth.bgr_to_nv12(left)
    data = capi.post_pack(q, capi.jpeg_encode_nv12(nv, w, h))
    # ---- the render tool, restated (no rclpy here): slicing :57-62, uint32 view :63-66, scale :29,73-75, depth :30-31,81-82
    infer_data_len = w * h * 4
    buf_infer, buf_jpeg = data[0:infer_data_len], data[infer_data_len:len(data) + 1]
    data_infer = np.ndarray(shape=(1, w * h), dtype=np.uint32, buffer=buf_infer)
    image_pre = data_infer.reshape((1, 1, 720, 1280)) * 2.60443857769133e-06
    image_pre = image_pre[-1] * 16 * 12
    Z = 527.1931762695312 * 119.89382172 / image_pre / 1000
    color = cv2.applyColorMap(cv2.convertScaleAbs(Z.squeeze(0), alpha=9), cv2.COLORMAP_JET)
    data_jpeg = np.ndarray(shape=(1, len(buf_jpeg)), dtype=np.uint8, buffer=buf_jpeg)
    im = cv2.imdecode(data_jpeg, cv2.IMREAD_ANYCOLOR)                                      # :95 - None here crashed the tool at :97
    assert im is not None and im.astype("uint8").shape == (720, 1280, 3)
    assert np.abs(image_pre[0] - disp).max() < 192 * 2.60443857769133e-06                  # disparity survives the wire format
    assert color.shape == (720, 1280, 3)
    assert _psnr(im, cv2.cvtColor(nv.reshape(h * 3 // 2, w), cv2.COLOR_YUV2BGR_NV12)) > 35.0
    d_ref, c_ref = pp.render_depth_colormap(q[:, 0], alpha=9.0)
    assert (c_ref[0] == color).all()
