"""GPU (B200): edge cases and full-size properties of the tcgen05 path (SNB_PREC_TC_F16X2), through the C ABI.

Small shapes are checked against the fp32 oracle (bar: mean EPE <= 1e-3 px, max <= 2e-2 px).  At BASELINE.json's
full sizes, where the CPU oracle takes minutes, the checks are size-independent properties: determinism, batch
invariance, range, and agreement of the tensor-core path with the library's exact-arithmetic fp32 path (which the
small-shape tests pin to the oracle)."""
import numpy as np
import pytest

from oracle import arch, prepost_ref as pp, synth, weights
from oracle.stereonet_ref import Oracle

pytestmark = pytest.mark.gpu
EPE_BAR, MAX_BAR = 1e-3, 2e-2


def _px(q):
    return q.astype(np.float64) * arch.OUT_SCALE * arch.OUT_NORM


def _s8(cfg, seed, n=1):
    if cfg.H % 2 or cfg.W % 2:        # no NV12 frame has odd sides: build the s8 tensor from an even-sized one, cropped
        big = arch.Config(cfg.H + cfg.H % 2, cfg.W + cfg.W % 2, cfg.K, cfg.D)
        return np.ascontiguousarray(_s8(big, seed, n)[:, :, :cfg.H, :cfg.W])
    out = []
    for i in range(n):
        frame = synth.frame(cfg.H, cfg.W, cfg.max_disp, seed=seed + i)
        out.append(pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(frame, cfg.H, 2 * cfg.W), cfg.W, cfg.H))
    return np.concatenate(out)


def _model(cfg, tc=True, **kw):
    from hobot_stereonet_b200 import Model, capi
    return Model(cfg.H, cfg.W, cfg.K, cfg.D, weights=weights.make_blob(cfg.K, seed=1234),
                 precision=capi.PREC_TC_F16X2 if tc else capi.PREC_FP32, **kw)


def _check_vs_oracle(cfg, s8, q):
    ref = Oracle(cfg, weights.generate(cfg.K, seed=1234)).forward_px(s8)
    err = np.abs(_px(q)[:, 0] - ref)
    assert err.mean() <= EPE_BAR and err.max() <= MAX_BAR, (err.mean(), err.max())
    return err


@pytest.mark.parametrize("H,W,K,D", [
    (32, 64, 3, 12),      # D > cost-volume width (8): hypotheses with shift >= width are all-zero slices
    (46, 154, 3, 6),      # KITTI-like ragged shape: pads to 48 x 160, crop on output
    (30, 34, 2, 5),       # tiny, odd multiples, K = 2 (no stride in layer2/3)
    (45, 77, 3, 4),       # odd height and width (KITTI's 375 rows): s8 entry point only, pads to 48 x 80
    (40, 1100, 2, 4),     # wide: 9 strips at full res, 3 at 1/4 res in the fused block / streaming kernels
    (272, 64, 3, 4),      # tall and narrow: many row chunks per column walk
])
def test_tc_edge_shapes_vs_oracle(built_lib, H, W, K, D):
    cfg = arch.Config(H, W, K, D)
    s8 = _s8(cfg, seed=300 + H)
    m = _model(cfg)
    q = m.infer(s8)
    assert (q == m.infer(s8)).all()
    m.close()
    _check_vs_oracle(cfg, s8, q)


def test_nv12_entry_rejects_odd_sizes(built_lib):
    from hobot_stereonet_b200 import SnbError, capi
    cfg = arch.Config(45, 77, 3, 4)
    m = _model(cfg)
    with pytest.raises(SnbError) as e:
        m.infer_nv12(np.zeros((1, 45 * 3 // 2, 2 * 77), np.uint8))
    assert e.value.code == capi.SNB_ERR_INVALID
    m.close()


def test_tc_batch_chunking_and_invariance(built_lib):
    cfg = arch.Config(48, 160, 3, 6)
    s8 = _s8(cfg, seed=500, n=5)
    m2 = _model(cfg, max_batch=2)
    q_all = m2.infer(s8)                       # 5 pairs through chunks of 2, 2, 1
    m1 = _model(cfg, max_batch=1)
    for i in range(5):
        assert (m1.infer(s8[i:i + 1]) == q_all[i:i + 1]).all()
    perm = [3, 0, 4, 1, 2]
    assert (m2.infer(np.ascontiguousarray(s8[perm])) == q_all[perm]).all()
    m1.close(); m2.close()
    _check_vs_oracle(cfg, s8, q_all)


def test_tc_unfused_and_tiled_paths_agree(built_lib):
    """The diagnostic flags select older kernels for the same layers; every combination must meet the same bar."""
    from hobot_stereonet_b200 import capi
    cfg = arch.Config(64, 160, 3, 8)
    s8 = _s8(cfg, seed=700)
    ref = Oracle(cfg, weights.generate(cfg.K, seed=1234)).forward_px(s8)
    for flags in (0, capi.FLAG_NO_FUSE, capi.FLAG_NO_STREAM, capi.FLAG_NO_FUSE | capi.FLAG_NO_STREAM, capi.FLAG_NO_GRAPH, capi.FLAG_NO_HBMCONV, capi.FLAG_NO_HEADFUSE):
        m = _model(cfg, flags=flags)
        err = np.abs(_px(m.infer(s8))[:, 0] - ref)
        m.close()
        assert err.mean() <= EPE_BAR and err.max() <= MAX_BAR, (flags, err.mean(), err.max())


def test_config3_shape_batch2_vs_oracle(built_lib):
    """BASELINE.json configs[2] shape (ZED 720x1280, D=48, K=3) at batch 2 of its 8."""
    cfg = arch.Config(720, 1280, 3, 48)
    s8 = _s8(cfg, seed=900, n=2)
    m = _model(cfg, max_batch=2)
    q = m.infer(s8)
    m.close()
    err = _check_vs_oracle(cfg, s8, q)
    print(f"config3 shape mean EPE {err.mean():.3e} px, max {err.max():.3e} px")


@pytest.mark.parametrize("H,W,D,B", [(540, 960, 192, 2), (375, 1242, 192, 2)])
def test_high_disparity_full_size_properties(built_lib, H, W, D, B):
    """BASELINE.json configs[3] / [4] shapes (D = 192): determinism, batch invariance, range, and the tensor-core path against
    the exact fp32 GPU path at the flat 1e-3 px bar."""
    cfg = arch.Config(H, W, 3, D)
    s8 = _s8(cfg, seed=1100 + H, n=B)
    m = _model(cfg, max_batch=B)
    q = m.infer(s8)
    assert (q == m.infer(s8)).all()                                   # deterministic
    assert (m.infer(s8[1:2]) == q[1:2]).all()                         # a pair does not see its batch neighbour
    m.close()
    assert q.shape == (B, 1, H, W) and q.min() >= 0 and _px(q).max() < 2 * cfg.max_disp
    # agreement with the exact-arithmetic fp32 CUDA path at the flat bar (both are checked against the CPU oracle at this size
    # in tests/test_gpu_d192.py)
    mf = _model(cfg, tc=False, max_batch=1)
    qf = mf.infer(s8[:1])
    mf.close()
    err = np.abs(_px(q[:1]) - _px(qf))
    print(f"{H}x{W} D={D}: tensor-core vs fp32 path mean {err.mean():.3e} px, max {err.max():.3e} px")
    assert err.mean() <= EPE_BAR and err.max() <= MAX_BAR


def test_gpu_depth_and_colormap_bit_exact(built_lib):
    """ParseTensor (parser.cpp:79-118) on the GPU, bit-exact against cv2 for alpha = 11 and 9."""
    cfg = arch.Config(64, 96, 3, 8)
    rng = np.random.default_rng(5)
    q = rng.integers(0, 400000, (2, 1, cfg.H, cfg.W), dtype=np.int64).astype(np.int32)
    q[0, 0, 0, :8] = [0, 1, 2, 3, 1000, 2 ** 31 - 1, 383962, 7]          # inf, huge, saturating and ordinary depths
    q[1, 0, 1, :4] = [5000, 50000, 500000, 5000000]
    m = _model(cfg)
    s8 = _s8(cfg, seed=77)
    qn = m.infer(s8)                                                       # and a real network output
    for alpha in (11.0, 9.0):
        for arr in (q, qn):
            depth, bgr = m.depth_color(arr, alpha)
            d_ref, c_ref = pp.render_depth_colormap(arr[:, 0], alpha=alpha)
            assert (depth.view(np.uint32) == d_ref.view(np.uint32)).all()      # bit-exact floats (inf included)
            assert (bgr == c_ref).all()
    m.close()
