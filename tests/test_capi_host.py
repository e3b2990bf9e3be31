"""CPU: the C-ABI library loads, exports every symbol include/snb200.h declares, its host-side
byte-format functions are bit-exact against the oracle, and compute entry points fail loudly
without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import prepost_ref as pp, synth, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "snb200.h")).read()
    names = re.findall(r"SNB_API\s+[\w\s\*]+?\b(snb_\w+)\s*\(", hdr)
    assert len(names) >= 20, names
    for n in names:
        assert hasattr(built_lib, n), f"libsnb200.so lacks {n}"


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hobot_stereonet_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "oracle/" not in src or f.endswith((".cu", ".cpp", ".cuh", ".h", ".py")) and \
                    not re.search(r"(open|dlopen|CDLL)\([^)]*oracle", src), f


@pytest.mark.parametrize("h,w", [(16, 24), (540, 960), (720, 1280), (376, 1242)])
def test_host_preprocess_bit_exact(built_lib, h, w):
    from hobot_stereonet_b200 import capi
    frame = synth.frame(h, w, 64, seed=h + w)
    l0, r0 = pp.split_side_by_side_nv12(frame, h, 2 * w)
    l1, r1 = capi.pre_split_nv12(frame, h, 2 * w)
    assert (l0 == l1).all() and (r0 == r1).all()
    for cc in (False, True):
        assert (pp.yuv420_to_yuv444(l0, w, h, cc) == capi.pre_yuv420_to_yuv444(l0, w, h, cc)).all()
        assert (pp.cvt_nv12_to_tensor_fast(l0, r0, w, h, cc) == capi.pre_cvt_nv12_to_tensor(l0, r0, w, h, cc)).all()


def test_host_preprocess_golden(built_lib):
    from hobot_stereonet_b200 import capi
    g = np.load(os.path.join(ROOT, "tests", "golden", "prepost_16x24.npz"))
    l, r = capi.pre_split_nv12(g["frame"], 16, 48)
    assert (capi.pre_cvt_nv12_to_tensor(l, r, 24, 16) == g["s8"]).all()
    assert (capi.pre_cvt_nv12_to_tensor(l, r, 24, 16, True) == g["s8_correct"]).all()


def test_host_quantize_all_bytes(built_lib):
    from hobot_stereonet_b200 import capi
    for x in range(256):
        v = float((np.float32(x) - np.float32(128.0)) / np.float32(128.0))
        assert capi.pre_quantize(v) == x - 128 == int(pp.quantize(v))
    assert capi.pre_quantize(5.0) == 127 and capi.pre_quantize(-5.0) == -128


def test_host_preprocess_rejects_bad_input(built_lib):
    from hobot_stereonet_b200 import capi
    buf = np.zeros(64, np.uint8)
    # preprocess.cpp:919-922: null inputs -> -1
    assert built_lib.snb_pre_cvt_nv12_to_tensor(None, buf.ctypes.data, 4, 4, 0, buf.ctypes.data) == capi.SNB_ERR_INVALID
    assert built_lib.snb_pre_split_nv12(buf.ctypes.data, 3, 8, buf.ctypes.data, buf.ctypes.data) == capi.SNB_ERR_INVALID


def test_host_pack_and_parse(built_lib):
    from hobot_stereonet_b200 import capi
    rng = np.random.default_rng(5)
    q = rng.integers(1, 400000, (1, 1, 36, 48)).astype(np.int32)
    jpeg = bytes(rng.integers(0, 256, 777, dtype=np.uint8))
    assert capi.post_pack(q, jpeg) == pp.pack_output(q, jpeg)
    assert capi.post_pack(q, b"") == pp.pack_output(q, b"")
    d = capi.post_parse_depth(q)
    assert (d == pp.parse_tensor_depth_f32(q)).all()
    assert np.allclose(d, pp.depth_m(q), rtol=1e-6)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu(built_lib):
    from hobot_stereonet_b200 import Model, SnbError, capi
    with pytest.raises(SnbError) as e:
        Model(64, 96, 3, 8, weights=weights.make_blob(3))
    assert e.value.code == capi.SNB_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_create_checks_model_file_first(built_lib):
    from hobot_stereonet_b200 import Model, SnbError, capi
    with pytest.raises(SnbError) as e:      # SetNodePara: "File is not exist! model_file: ..." (node.cpp:131-134)
        Model(64, 96, 3, 8, model_file="/nonexistent/hobot_stereonet.hbm")
    assert e.value.code == capi.SNB_ERR_MODEL and "File is not exist" in str(e.value)
    with pytest.raises(SnbError) as e:
        Model(0, 96, 3, 8, weights=b"x" * 64)
    assert e.value.code == capi.SNB_ERR_INVALID
    for K in (1, 5):                        # the backbone's stride schedule exists for 1/4, 1/8 and 1/16 volumes only
        with pytest.raises(SnbError) as e:
            Model(64, 96, K, 8, weights=b"x" * 64)
        assert e.value.code == capi.SNB_ERR_INVALID and "2<=K<=4" in str(e.value)


def test_synthesized_blob_matches_oracle_layer_table(built_lib):
    """The library's own layer table (weights_host.cpp) and the oracle's (oracle/arch.py) agree."""
    from hobot_stereonet_b200 import capi
    from oracle import arch
    for K in (2, 3, 4):
        k2, w = weights.from_blob(capi.synthesize_weights(K, 7))
        ref = weights.generate(K)
        assert k2 == K and list(w.keys()) == list(ref.keys())
        assert all(w[n].shape == ref[n].shape for n in w)
        for spec in arch.conv_specs(K):
            fan_in = spec.cin * int(np.prod(spec.k))
            std = w[spec.name + ".weight"].std()
            assert abs(std / (spec.gain * np.sqrt(2.0 / fan_in)) - 1) < 0.25, spec.name
    assert capi.synthesize_weights(3, 7) == capi.synthesize_weights(3, 7) != capi.synthesize_weights(3, 8)
