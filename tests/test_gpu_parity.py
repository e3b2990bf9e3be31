"""GPU (B200): parity of the CUDA path, called through the C ABI, against the fp32 oracle and the
committed golden fixtures.  Tolerance: mean end-point error <= 1e-3 px (BASELINE.json north_star);
integer pre-process paths bit-exact."""
import os
import threading

import numpy as np
import pytest

from oracle import arch, prepost_ref as pp, synth, weights
from oracle.stereonet_ref import Oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPE_BAR = 1e-3          # px, mean |disp_gpu - disp_oracle| over valid pixels
MAX_BAR = 2e-2          # px, worst pixel (one s32 step is 5e-4 px)
CASES = ["net_64x96_k3_d8", "net_50x70_k2_d6", "net_64x128_k4_d4", "net_40x48_k3_d12"]


def _px(q):
    return q.astype(np.float64) * arch.OUT_SCALE * arch.OUT_NORM


def _model(H, W, K, D, **kw):
    from hobot_stereonet_b200 import Model
    return Model(H, W, K, D, weights=weights.make_blob(K, seed=kw.pop("seed", 1234)), **kw)


@pytest.mark.parametrize("name", CASES)
def test_golden_fixture(built_lib, name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    H, W, K, D, B = [int(v) for v in g["cfg"]]
    m = _model(H, W, K, D, max_batch=B)
    q = m.infer(g["s8"])
    assert q.shape == (B, 1, H, W) and q.dtype == np.int32
    err = np.abs(_px(q) - _px(g["q"]))
    assert err.mean() <= EPE_BAR and err.max() <= MAX_BAR, (err.mean(), err.max())
    # raw camera frames through the GPU pre-process give the identical tensor -> identical output
    q2 = m.infer_nv12(g["frames"].reshape(B, H * 3 // 2, 2 * W))
    assert (q == q2).all()
    m.close()


def test_stage_parity(built_lib):
    from hobot_stereonet_b200 import capi
    g = np.load(os.path.join(ROOT, "tests", "golden", "net_64x96_k3_d8.npz"))
    H, W, K, D, B = [int(v) for v in g["cfg"]]
    cfg = arch.Config(H, W, K, D)
    m = _model(H, W, K, D, flags=capi.FLAG_KEEP_STAGES)
    m.infer(g["s8"])
    dump = {}
    Oracle(cfg, weights.generate(K, seed=1234)).forward_norm(g["s8"], dump)
    worst = {}
    for name in ["firstconv", "layer1", "layer2", "layer3", "layer4", "gwc", "cat", "volume", "filter0", "filter4",
                 "cost", "disp0", "refine0.feat", "disp1", "refine2.feat", "disp3"]:
        ref = dump[name].numpy()
        got = m.debug_read(name)
        if got.ndim == ref.ndim + 1:
            got = got[:, 0]
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        scale = max(1.0, float(np.abs(ref).max()))
        worst[name] = float(np.abs(got - ref).max()) / scale
        assert worst[name] < 2e-4, (name, worst)
    m.close()


def test_batch_chunking_and_invariance(built_lib):
    H, W, K, D = 48, 64, 3, 6
    cfg = arch.Config(H, W, K, D)
    frames = np.stack([synth.frame(H, W, cfg.max_disp, seed=40 + i) for i in range(5)])
    s8 = np.concatenate([pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(f, H, 2 * W), W, H) for f in frames])
    m2 = _model(H, W, K, D, max_batch=2)
    q_all = m2.infer(s8)                       # 5 pairs through chunks of 2,2,1
    m1 = _model(H, W, K, D, max_batch=1)
    for i in range(5):
        assert (m1.infer(s8[i:i + 1]) == q_all[i:i + 1]).all()
    # a pair's result does not depend on its batch neighbours
    perm = [3, 0, 4, 1, 2]
    assert (m2.infer(np.ascontiguousarray(s8[perm])) == q_all[perm]).all()
    m1.close(); m2.close()


def test_async_api(built_lib):
    H, W, K, D = 48, 64, 3, 6
    cfg = arch.Config(H, W, K, D)
    s8 = [pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(synth.frame(H, W, cfg.max_disp, seed=60 + i), H, 2 * W), W, H)
          for i in range(6)]
    m = _model(H, W, K, D, task_num=4)
    want = [m.infer(x) for x in s8]
    outs = [np.zeros((1, 1, H, W), np.int32) for _ in s8]
    seen, lock = [], threading.Lock()

    def done(i):
        def f(status, stat):
            with lock:
                seen.append((i, status, threading.get_ident()))
        return f

    for i, x in enumerate(s8):
        m.infer_async(x, outs[i], done(i))     # 6 tasks through 4 slots: blocks, never drops
    m.wait_all()
    assert sorted(i for i, _, _ in seen) == list(range(6)) and all(s == 0 for _, s, _ in seen)
    assert all(t != threading.get_ident() for _, _, t in seen)      # library-owned PostProcess thread
    assert all((a == b).all() for a, b in zip(outs, want))
    m.close()


@pytest.mark.parametrize("flags_name", ["coalesce", "no_coalesce"])
def test_async_calls_coalesced_into_batched_passes(built_lib, flags_name):
    """max_batch > 1: the worker merges queued snb_infer_async calls into passes of up to max_batch pairs (two passes in
    flight).  Per-call semantics must not change: every call gets ITS result (bit-identical to the synchronous call,
    whatever pass and position it lands in), its own callback, in submission order."""
    from hobot_stereonet_b200 import capi
    H, W, K, D = 48, 64, 3, 6
    cfg = arch.Config(H, W, K, D)
    n = 23
    s8 = [pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(synth.frame(H, W, cfg.max_disp, seed=160 + i), H, 2 * W), W, H)
          for i in range(n)]
    flags = capi.FLAG_NO_COALESCE if flags_name == "no_coalesce" else 0
    m = _model(H, W, K, D, task_num=4, max_batch=3, flags=flags)
    want = [m.infer(x) for x in s8]
    outs = [np.zeros((1, 1, H, W), np.int32) for _ in s8]
    seen, lock = [], threading.Lock()

    def done(i):
        def f(status, stat):
            with lock:
                seen.append((i, status))
        return f

    for rep in range(2):                       # second round: slots and graphs of every batch size already exist
        seen.clear()
        for o in outs:
            o[:] = 0
        for i, x in enumerate(s8):
            m.infer_async(x, outs[i], done(i))
            if i == 9:
                m.wait_all()                   # drain in the middle: the worker goes idle and restarts
        m.wait_all()
        assert [i for i, _ in seen] == list(range(n)) and all(s == 0 for _, s in seen)
        assert all((a == b).all() for a, b in zip(outs, want))
    # a call that is itself a batch of 2 coalesces with a single (2 + 1 <= 3) and keeps its layout
    pair = np.concatenate([s8[0], s8[1]])
    out2, out1 = np.zeros((2, 1, H, W), np.int32), np.zeros((1, 1, H, W), np.int32)
    m.infer_async(pair, out2)
    m.infer_async(s8[2], out1)
    m.wait_all()
    assert (out2[0:1] == want[0]).all() and (out2[1:2] == want[1]).all() and (out1 == want[2]).all()
    m.close()


def test_io_properties_and_errors(built_lib):
    from hobot_stereonet_b200 import Model, SnbError, capi
    m = _model(64, 96, 3, 8)
    i, o = m.io_props()
    assert list(i.valid_shape) == [1, 6, 64, 96] and i.tensor_type == capi.TENSOR_S8 and i.tensor_layout == capi.LAYOUT_NCHW
    assert list(o.valid_shape) == [1, 1, 64, 96] and o.tensor_type == capi.TENSOR_S32
    assert abs(i.scale - 1 / 128) < 1e-12 and abs(o.scale - 2.60443857769133e-06) < 1e-12
    assert i.mem_size == 6 * 64 * 96 and o.mem_size == 4 * 64 * 96
    assert m.model_input_size() == (96, 64)
    assert capi.lib().snb_infer(m._h, None, None, 1) == capi.SNB_ERR_INVALID
    m.close()
    with pytest.raises(SnbError) as e:       # blob generated for another K
        Model(64, 96, 3, 8, weights=weights.make_blob(2))
    assert e.value.code == capi.SNB_ERR_MODEL


def test_set_weights_swaps_model(built_lib):
    g = np.load(os.path.join(ROOT, "tests", "golden", "net_64x96_k3_d8.npz"))
    cfg = arch.Config(64, 96, 3, 8)
    m = _model(64, 96, 3, 8)
    q0 = m.infer(g["s8"])
    m.set_weights(weights.make_blob(3, seed=77))
    q1 = m.infer(g["s8"])
    assert not (q0 == q1).all()
    ref = Oracle(cfg, weights.generate(3, seed=77)).forward_px(g["s8"])
    assert np.abs(_px(q1)[:, 0] - ref).mean() <= EPE_BAR
    m.close()


def test_config2_full_size_epe(built_lib):
    """BASELINE.json configs[1]: 540x960, K=3, D=24, batch 1."""
    cfg = arch.Config(540, 960, 3, 24)
    frame = synth.frame(cfg.H, cfg.W, cfg.max_disp, seed=1235)
    s8 = pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(frame, cfg.H, 2 * cfg.W), cfg.W, cfg.H)
    m = _model(cfg.H, cfg.W, cfg.K, cfg.D)
    q = m.infer(s8)
    q_nv12 = m.infer_nv12(frame.reshape(1, cfg.H * 3 // 2, 2 * cfg.W))
    m.close()
    assert (q == q_nv12).all()
    ref = Oracle(cfg, weights.generate(cfg.K, seed=1234)).forward_px(s8)
    err = np.abs(_px(q)[:, 0] - ref)
    print(f"config2 mean EPE {err.mean():.3e} px, max {err.max():.3e} px")
    assert err.mean() <= EPE_BAR and err.max() <= MAX_BAR
    # size-independent properties at full size: non-negative (final ReLU), bounded, finite
    assert q.min() >= 0 and _px(q).max() < 2 * cfg.max_disp
