"""Multi-GPU host logic without a GPU: the C++ dispatch arithmetic (host/dispatch.h, unit test lib/test_dispatch), its C ABI
twin snb_shard_range against the Python shard_range used under torchrun, and the weight-blob validation that guards
snb_create / snb_set_weights / snb_pool_create."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_dispatch_unit_test(built_lib):
    exe = os.path.join(ROOT, "hobot_stereonet_b200", "lib", "test_dispatch")
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "dispatch ok" in r.stdout, r.stderr


def test_c_shard_range_matches_python(built_lib):
    from hobot_stereonet_b200 import capi
    from hobot_stereonet_b200.shard import shard_range
    for world in range(1, 10):
        for n in (0, 1, 7, 8, 9, 32, 63, 64, 65):
            for rank in range(world):
                assert capi.shard_range(n, world, rank) == shard_range(n, world, rank)
    assert capi.shard_range(32, 8, 0) == (0, 4) and capi.shard_range(64, 8, 7) == (56, 64)      # BASELINE configs 4 / 5
    with pytest.raises(capi.SnbError):
        capi.shard_range(4, 2, 2)


def test_pool_without_gpu_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from hobot_stereonet_b200 import Pool, SnbError, capi
    with pytest.raises(SnbError) as e:
        Pool(64, 96, 3, 8, weights=weights.make_blob(3))
    assert e.value.code == capi.SNB_ERR_CUDA and "no CUDA device" in str(e.value)


def _mutate(blob: bytes, name: str, field: str, value) -> bytes:
    """Rewrite one field of the table entry of tensor `name`."""
    b = bytearray(blob)
    n = struct.unpack_from("<I", b, 16)[0]
    for i in range(n):
        off = 24 + i * 104
        if bytes(b[off:off + 64]).rstrip(b"\0").decode() == name:
            pos = {"ndim": 64, "dim0": 68, "dim1": 72, "offset": 88, "nbytes": 96}[field]
            struct.pack_into("<Q" if field in ("offset", "nbytes") else "<I", b, off + pos, value)
            return bytes(b)
    raise KeyError(name)


def test_weight_blob_validation(built_lib):
    """A model_file is untrusted input: every way the table can lie is an SNB_ERR_MODEL, not an out-of-bounds read."""
    from hobot_stereonet_b200 import SnbError, capi
    blob = weights.make_blob(3, seed=1)
    capi.weights_validate(blob)
    capi.weights_validate(blob, 3)
    capi.weights_validate(capi.synthesize_weights(4, 7), 4)
    bad = {
        "wrong K": (blob, 4),
        "magic": (b"XNB2WGT1" + blob[8:], 0),
        "truncated table": (blob[:500], 0),
        "truncated data": (blob[:len(blob) // 2], 0),
        "offset overflow": (_mutate(blob, "backbone.firstconv.1.weight", "offset", 2 ** 64 - 64), 0),
        "nbytes overflow": (_mutate(blob, "backbone.firstconv.1.weight", "nbytes", 2 ** 64 - 4), 0),
        "nbytes != prod(dims)": (_mutate(blob, "backbone.firstconv.1.weight", "nbytes", 32 * 32 * 9 * 4 - 4), 0),
        "ndim 2": (_mutate(blob, "backbone.firstconv.1.weight", "ndim", 2), 0),
        "ndim 9": (_mutate(blob, "backbone.firstconv.1.weight", "ndim", 9), 0),
        "cout of another layer": (_mutate(_mutate(blob, "backbone.firstconv.1.weight", "dim0", 16), "backbone.firstconv.1.weight", "nbytes", 16 * 32 * 9 * 4), 0),
        "short bias": (_mutate(_mutate(blob, "backbone.layer2.3.conv_a.bias", "dim0", 8), "backbone.layer2.3.conv_a.bias", "nbytes", 32), 0),
        "zero dim": (_mutate(blob, "head.filter.0.weight", "dim1", 0), 0),
    }
    for why, (b, K) in bad.items():
        with pytest.raises(SnbError) as e:
            capi.weights_validate(b, K)
        assert e.value.code == capi.SNB_ERR_MODEL, why
    # a missing layer
    t = weights.generate(3, seed=1)
    del t["head.refine.2.conv_out.weight"]
    with pytest.raises(SnbError) as e:
        capi.weights_validate(weights.to_blob(t, 3))
    assert "lacks head.refine.2.conv_out" in str(e.value)
    # a transposed layer (same byte count, wrong shape)
    t = weights.generate(3, seed=1)
    t["backbone.layer2.0.conv_a.weight"] = np.ascontiguousarray(t["backbone.layer2.0.conv_a.weight"].transpose(1, 0, 2, 3))
    with pytest.raises(SnbError):
        capi.weights_validate(weights.to_blob(t, 3))
