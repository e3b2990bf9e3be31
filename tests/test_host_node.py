"""The C++ host-side mirror of the reference node (hobot_stereonet_b200/host/): StereonetNode over the
dnn_node stand-in over the C ABI, driven through the `stereonet_infer` executable.
CPU part: builds, fails like the reference when the model file is missing, tensor memory allocator works.
GPU part: the published payload ([s32 x H*W] || [JPEG], stereonet_node.cpp:1033-1049) is bit-identical to
what the ctypes binding returns for the same frames, bad frames are dropped (stereonet_node.cpp:672-690)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import arch, prepost_ref as pp, synth, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "hobot_stereonet_b200", "lib", "stereonet_infer")


def test_cli_missing_model_fails_like_reference(built_lib):
    assert os.path.exists(EXE)
    r = subprocess.run([EXE, "--model_file", "/nonexistent/hobot_stereonet.hbm", "--frames", "x", "--out", "y"],
                       capture_output=True, text=True)
    assert r.returncode == 1
    assert "File is not exist! model_file: /nonexistent/hobot_stereonet.hbm" in r.stderr      # stereonet_node.cpp:131-134
    assert "Node init fail!" in r.stderr                                                       # stereonet_node.cpp:46
    # the four reference parameters are echoed with their defaults (stereonet_node.cpp:27-41)
    assert "sub_hbmem_topic_name: hbmem_stereo_img" in r.stderr and "ros_img_topic_name: /stereonet_node_output" in r.stderr


def test_ros2_shell_compiles_and_links_against_stub_rclcpp(built_lib, tmp_path):
    """the ROS 2 shell (hobot_stereonet_b200/ros2) cannot be built here (no rclcpp in the image);
    it is compiled and LINKED against minimal stand-in headers (tests/ros_stubs) with the real host library, so the calls it
    makes into StereonetNode and the message fields it forwards (stereonet_node.cpp:27-35,108-118,657,1064) stay in sync."""
    src = os.path.join(ROOT, "hobot_stereonet_b200", "ros2", "src", "stereonet_ros_node.cpp")
    lib = os.path.join(ROOT, "hobot_stereonet_b200", "lib")
    exe = str(tmp_path / "ros_shell")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "tests", "ros_stubs"),
                        "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "hobot_stereonet_b200", "host"), src, "-o", exe,
                        "-L" + lib, "-lstereonet_host", "-lsnb200", "-Wl,-rpath," + lib, "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # the default model_file does not exist here: the shell must fail like the reference node (log + shutdown), not crash
    r = subprocess.run([exe], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0 and "Node init fail!" in r.stderr
    # same package / executable names and launch arguments as the reference (hobot_stereonet.launch.py:35-49)
    launch = open(os.path.join(ROOT, "hobot_stereonet_b200", "ros2", "launch", "hobot_stereonet.launch.py")).read()
    for needle in ('package="hobot_stereonet"', 'executable="hobot_stereonet"', '"config_file"', '"model_file"'):
        assert needle in launch


def test_sys_alloc_roundtrip(built_lib):
    p = C.c_void_p()
    assert built_lib.snb_sys_alloc(C.byref(p), 4096) == 0 and p.value and p.value % 64 == 0
    C.memset(p, 0x5A, 4096)
    assert C.string_at(p.value + 4095, 1) == b"\x5a"
    built_lib.snb_sys_free(p)
    assert built_lib.snb_sys_alloc(C.byref(p), 0) < 0


@pytest.mark.gpu
def test_cli_payload_matches_binding(built_lib, tmp_path):
    from hobot_stereonet_b200 import Model, capi
    H, W, K, D, N = 64, 96, 3, 8, 6
    cfg = arch.Config(H, W, K, D)
    blob = weights.make_blob(K, seed=1234)
    (tmp_path / "model.bin").write_bytes(blob)
    frames = np.stack([synth.frame(H, W, cfg.max_disp, seed=300 + i) for i in range(N)])
    (tmp_path / "frames.nv12").write_bytes(frames.tobytes())
    r = subprocess.run([EXE, "--model_file", str(tmp_path / "model.bin"), "--frames", str(tmp_path / "frames.nv12"),
                        "--out", str(tmp_path / "out.bin"), "--model_in_h", str(H), "--model_in_w", str(W),
                        "--K", str(K), "--D", str(D), "--precision", "tc"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert f"fed {N} frame(s), published {N}, dropped 0" in r.stderr
    raw = np.frombuffer((tmp_path / "out.bin").read_bytes(), np.uint8)
    m = Model(H, W, K, D, weights=blob, precision=capi.PREC_TC_F16X2)
    pos, seen = 0, set()
    for i in range(N):                       # async: arrival order is the task order (one worker), ids still checked
        hdr = raw[pos:pos + 16].view(np.uint32)
        idx, step = int(hdr[0]), int(hdr[3]); seen.add(idx)
        assert (int(hdr[1]), int(hdr[2])) == (H, W) and step > H * W * 4             # height, width, step = len(data)
        data = raw[pos + 16:pos + 16 + step]
        pos += 16 + step
        q = data[:H * W * 4].view(np.int32).reshape(1, 1, H, W)
        s8 = pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(frames[idx], H, 2 * W), W, H)
        assert (q == m.infer(s8)).all()
        # the JPEG half: the left view, readable by the render tool's cv2.imdecode (publisher_member_function.py:93-95)
        import cv2
        im = cv2.imdecode(data[H * W * 4:], cv2.IMREAD_ANYCOLOR)
        left = pp.split_side_by_side_nv12(frames[idx], H, 2 * W)[0]
        ref = cv2.cvtColor(left.reshape(H * 3 // 2, W), cv2.COLOR_YUV2BGR_NV12)
        assert im is not None and im.shape == ref.shape
        assert 10 * np.log10(255.0 ** 2 / np.mean((im.astype(np.float64) - ref) ** 2)) > 30.0
    assert pos == raw.size and seen == set(range(N))
    m.close()
    # the reference's host pre-process path (preprocess=cpu) and a JPEG-less payload give the same s32 bytes
    r = subprocess.run([EXE, "--model_file", str(tmp_path / "model.bin"), "--frames", str(tmp_path / "frames.nv12"),
                        "--out", str(tmp_path / "out_cpu.bin"), "--model_in_h", str(H), "--model_in_w", str(W),
                        "--K", str(K), "--D", str(D), "--precision", "tc", "--preprocess", "cpu", "--jpeg", "off"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw2 = np.frombuffer((tmp_path / "out_cpu.bin").read_bytes(), np.uint8)
    rec = 16 + H * W * 4
    assert raw2.size == N * rec
    first = {int(raw[0:16].view(np.uint32)[0]): raw[16:16 + H * W * 4]}
    for i in range(N):
        hdr = raw2[i * rec:i * rec + 16].view(np.uint32)
        assert int(hdr[3]) == H * W * 4
        if int(hdr[0]) in first:
            assert (raw2[i * rec + 16:(i + 1) * rec] == first[int(hdr[0])]).all()


@pytest.mark.gpu
def test_cli_multi_device(built_lib, tmp_path):
    """--devices 0,1,...: one replica per GPU behind the same node (snb_pool_*); every frame is published exactly once."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    H, W, K, D, N = 64, 96, 3, 8, 24
    (tmp_path / "model.bin").write_bytes(weights.make_blob(K, seed=1234))
    frames = np.stack([synth.frame(H, W, 64, seed=300 + i) for i in range(N)])
    (tmp_path / "frames.nv12").write_bytes(frames.tobytes())
    r = subprocess.run([EXE, "--model_file", str(tmp_path / "model.bin"), "--frames", str(tmp_path / "frames.nv12"),
                        "--out", str(tmp_path / "out.bin"), "--model_in_h", str(H), "--model_in_w", str(W), "--K", str(K), "--D", str(D),
                        "--devices", ",".join(str(i) for i in range(ndev)), "--jpeg", "off"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert f"fed {N} frame(s), published {N}, dropped 0" in r.stderr


@pytest.mark.gpu
def test_cli_drops_bad_frames(built_lib, tmp_path):
    H, W, K, D = 64, 96, 3, 8
    (tmp_path / "model.bin").write_bytes(weights.make_blob(K, seed=1234))
    frame = synth.frame(H, W, 64, seed=1)
    (tmp_path / "frames.nv12").write_bytes(frame.tobytes())
    base = [EXE, "--model_file", str(tmp_path / "model.bin"), "--frames", str(tmp_path / "frames.nv12"), "--out",
            str(tmp_path / "out.bin"), "--model_in_h", str(H), "--model_in_w", str(W), "--K", str(K), "--D", str(D)]
    r = subprocess.run(base + ["--encoding", "bgr8"], capture_output=True, text=True)
    assert r.returncode == 0 and "Only support nv12 img encoding" in r.stderr and "published 0, dropped 1" in r.stderr
