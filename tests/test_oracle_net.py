"""CPU: the oracle network reproduces the stage shapes the reference's .hbm tensor table records
for the deployed instance, the weight blob round-trips, and the
committed golden fixtures are reproducible."""
import os

import numpy as np
import pytest
import torch

from oracle import arch, weights
from oracle.stereonet_ref import Oracle, quant_multiplier

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_deployed_stage_shapes_match_hbm_table():
    cfg = arch.Config(720, 1280, 4, 12)
    assert (cfg.Hp, cfg.Wp, cfg.h, cfg.w, cfg.max_disp) == (720, 1280, 45, 80, 192)
    o = Oracle(cfg, weights.generate(4))
    o.w = {k: v.to("meta") for k, v in o.w.items()}        # shapes only: no arithmetic
    dump = {}
    out = o.forward_float(torch.empty(1, 6, 720, 1280, device="meta"), dump)
    shp = {k: tuple(v.shape) for k, v in dump.items()}
    assert shp["firstconv"] == (2, 32, 180, 320)            # hbm rec [24] 180x320x32
    assert shp["layer1"] == (2, 32, 180, 320)               # rec [19]
    assert shp["layer2"] == (2, 64, 90, 160)
    assert shp["layer3"] == (2, 128, 45, 80) and shp["layer4"] == (2, 128, 45, 80)
    assert shp["gwc"] == (2, 256, 45, 80)                   # rec [20] 45x80x256
    assert shp["cat"] == (2, 16, 45, 80)                    # rec [32,35] 45x80x16
    assert shp["volume"] == (1, 64, 12, 45, 80)             # rec [33] 1x64x540x80, 540 = 12*45
    assert shp["filter4"] == (1, 32, 12, 45, 80)
    assert shp["cost"] == (1, 12, 45, 80) and shp["disp0"] == (1, 45, 80)
    assert [shp[f"disp{i}"] for i in (1, 2, 3, 4)] == [(1, 90, 160), (1, 180, 320), (1, 360, 640), (1, 720, 1280)]
    assert tuple(out.shape) == (1, 720, 1280)               # rec [1] 1x720x1280x1


def test_param_count_and_macs_match_survey():
    n = sum(v.size for k, v in weights.generate(4).items() if k.endswith(".weight"))
    assert abs(n - 3.85e6) / 3.85e6 < 0.05                  # ~3.85 M unique params
    m = arch.macs_per_pair(arch.Config(720, 1280, 4, 12))
    assert abs(m["total"] / 1e9 - 204.4) < 1.0              # SURVEY §8d deployed shape
    m2 = arch.macs_per_pair(arch.Config(540, 960, 3, 24))
    assert abs(m2["total"] / 1e9 - 168.1) < 1.0


def test_blob_roundtrip():
    w = weights.generate(2, seed=9)
    K, w2 = weights.from_blob(weights.to_blob(w, 2))
    assert K == 2 and w.keys() == w2.keys()
    assert all((w[k] == w2[k]).all() for k in w)
    with pytest.raises(ValueError):
        weights.from_blob(b"notablob" + b"\0" * 64)


def test_quant_multiplier_decodes_to_pixels():
    cfg = arch.Config(540, 960, 3, 24)
    q = np.rint(np.float32(0.5) * quant_multiplier(cfg))
    assert abs(q * arch.OUT_SCALE * 192 - 0.5 * cfg.max_disp) < 1e-3


@pytest.mark.parametrize("name", ["net_64x96_k3_d8", "net_50x70_k2_d6", "net_64x128_k4_d4", "net_40x48_k3_d12"])
def test_golden_fixture_reproducible(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    H, W, K, D, B = [int(v) for v in g["cfg"]]
    cfg = arch.Config(H, W, K, D)
    o = Oracle(cfg, weights.generate(K, seed=1234))
    dn = o.forward_norm(g["s8"]).numpy()
    assert np.abs(dn - g["dn"]).max() * cfg.max_disp < 1e-4      # px; thread-count dependent summation only
    assert np.abs(o.forward_s32(g["s8"]).astype(np.int64) - g["q"]).max() <= 1


def test_cost_volume_definition():
    cfg = arch.Config(16, 32, 2, 4)
    o = Oracle(cfg, weights.generate(2))
    g = torch.Generator().manual_seed(0)
    gl, gr = torch.randn(1, 256, 4, 8, generator=g), torch.randn(1, 256, 4, 8, generator=g)
    cl, cr = torch.randn(1, 16, 4, 8, generator=g), torch.randn(1, 16, 4, 8, generator=g)
    v = o.cost_volume(gl, gr, cl, cr)
    assert v.shape == (1, 64, 4, 4, 8)
    d, y, x = 2, 1, 5
    assert torch.equal(v[0, :16, d, y, x], cl[0, :, y, x]) and torch.equal(v[0, 16:32, d, y, x], cr[0, :, y, x - d])
    want = (gl[0, 8 * 3:8 * 4, y, x] * gr[0, 8 * 3:8 * 4, y, x - d]).mean()
    assert torch.allclose(v[0, 32 + 3, d, y, x], want)
    assert (v[:, :, 2, :, :2] == 0).all()                    # zero where x < d
