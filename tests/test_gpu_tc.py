"""GPU (B200): the tcgen05 path (SNB_PREC_TC_F16X2: split-fp16 operands, fp32 accumulation in TMEM)
against the fp32 oracle.  Same bar as the fp32 path: mean EPE <= 1e-3 px."""
import os

import numpy as np
import pytest

from oracle import arch, prepost_ref as pp, synth, weights
from oracle.stereonet_ref import Oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPE_BAR, MAX_BAR = 1e-3, 2e-2
CASES = ["net_64x96_k3_d8", "net_50x70_k2_d6", "net_64x128_k4_d4", "net_40x48_k3_d12"]


def _px(q):
    return q.astype(np.float64) * arch.OUT_SCALE * arch.OUT_NORM


def _model(H, W, K, D, **kw):
    from hobot_stereonet_b200 import Model, capi
    return Model(H, W, K, D, weights=weights.make_blob(K, seed=1234), precision=capi.PREC_TC_F16X2, **kw)


def test_tc_stage_parity(built_lib):
    from hobot_stereonet_b200 import capi
    g = np.load(os.path.join(ROOT, "tests", "golden", "net_64x96_k3_d8.npz"))
    H, W, K, D, B = [int(v) for v in g["cfg"]]
    m = _model(H, W, K, D, flags=capi.FLAG_KEEP_STAGES)
    m.infer(g["s8"])
    dump = {}
    Oracle(arch.Config(H, W, K, D), weights.generate(K, seed=1234)).forward_norm(g["s8"], dump)
    report = []
    for name in ["firstconv", "layer1", "layer2", "layer3", "layer4", "gwc", "cat", "volume", "filter0", "filter4",
                 "cost", "disp0", "refine0.feat", "disp1", "refine2.feat", "disp3"]:
        ref = dump[name].numpy()
        got = m.debug_read(name)
        if got.ndim == ref.ndim + 1:
            got = got[:, 0]
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        rel = float(np.abs(got - ref).max()) / max(1.0, float(np.abs(ref).max()))
        report.append((name, rel))
    print(report)
    bad = [(n, r) for n, r in report if not r < 5e-4]
    assert not bad, report
    m.close()


@pytest.mark.parametrize("name", CASES)
def test_tc_golden_fixture(built_lib, name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    H, W, K, D, B = [int(v) for v in g["cfg"]]
    m = _model(H, W, K, D, max_batch=B)
    q = m.infer(g["s8"])
    err = np.abs(_px(q) - _px(g["q"]))
    assert err.mean() <= EPE_BAR and err.max() <= MAX_BAR, (err.mean(), err.max())
    assert (q == m.infer_nv12(g["frames"].reshape(B, H * 3 // 2, 2 * W))).all()
    assert (q == m.infer(g["s8"])).all()          # deterministic: no atomics, fixed tile order
    m.close()


def test_tc_config2_full_size_epe(built_lib):
    cfg = arch.Config(540, 960, 3, 24)
    frame = synth.frame(cfg.H, cfg.W, cfg.max_disp, seed=1235)
    s8 = pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(frame, cfg.H, 2 * cfg.W), cfg.W, cfg.H)
    m = _model(cfg.H, cfg.W, cfg.K, cfg.D)
    q = m.infer(s8)
    m.close()
    ref = Oracle(cfg, weights.generate(cfg.K, seed=1234)).forward_px(s8)
    err = np.abs(_px(q)[:, 0] - ref)
    print(f"tc config2 mean EPE {err.mean():.3e} px, max {err.max():.3e} px")
    assert err.mean() <= EPE_BAR and err.max() <= MAX_BAR


def test_tc_deployed_shape_k4(built_lib):
    """The reference's deployed instance: 720x1280, K=4, D=12 (hbm tensor table), batch 1."""
    cfg = arch.Config(720, 1280, 4, 12)
    frame = synth.frame(cfg.H, cfg.W, cfg.max_disp, seed=77)
    s8 = pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(frame, cfg.H, 2 * cfg.W), cfg.W, cfg.H)
    m = _model(cfg.H, cfg.W, cfg.K, cfg.D)
    q = m.infer(s8)
    m.close()
    ref = Oracle(cfg, weights.generate(cfg.K, seed=1234)).forward_px(s8)
    err = np.abs(_px(q)[:, 0] - ref)
    print(f"tc deployed-shape mean EPE {err.mean():.3e} px, max {err.max():.3e} px")
    assert err.mean() <= EPE_BAR and err.max() <= MAX_BAR
