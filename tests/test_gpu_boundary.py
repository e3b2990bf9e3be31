"""GPU (B200): the drop-in boundary beyond the plain Run() call, through the C ABI.

  * byte-exact GPU pre-process: snb_pre_nv12_gpu == oracle/prepost_ref (quirk and correct_chroma),
    == the committed golden bytes, == the host C function, and the image tensor the network reads == s8 / 128;
  * snb_infer_nv12_async: same results as the synchronous calls, callbacks in submission order, s8 and NV12 calls mixed;
  * a done-callback that resubmits while every task slot is taken (the reference's PostProcess could call Run);
  * snb_set_weights is transactional for bad blobs; a caller-stream snb_infer_device pass is ordered before the next pass;
  * snb_pool_*: one replica per GPU, weights by one NCCL broadcast (needs >= 2 GPUs for the collective itself).
"""
import os
import threading

import numpy as np
import pytest

from oracle import arch, prepost_ref as pp, synth, weights

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frames(H, W, n, seed, max_disp=64):
    return np.stack([synth.frame(H, W, max_disp, seed=seed + i).reshape(H * 3 // 2, 2 * W) for i in range(n)])


def _s8(frames, H, W, correct=False):
    return np.concatenate([pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(f.ravel(), H, 2 * W), W, H, correct_chroma=correct)
                           for f in frames])


def _model(H, W, K=3, D=8, seed=1234, **kw):
    from hobot_stereonet_b200 import Model, capi
    kw.setdefault("precision", capi.PREC_TC_F16X2)
    return Model(H, W, K, D, weights=weights.make_blob(K, seed=seed), **kw)


@pytest.mark.parametrize("H,W", [(16, 24), (540, 960), (720, 1280)])
@pytest.mark.parametrize("correct", [False, True])
def test_gpu_preprocess_byte_exact(built_lib, H, W, correct):
    """preprocess.h:128-155 (I420-style indexing of NV12 chroma, the quirk) + preprocess.cpp:999-1040 on the GPU."""
    from hobot_stereonet_b200 import capi
    if (H, W) == (16, 24):
        g = np.load(os.path.join(ROOT, "tests", "golden", "prepost_16x24.npz"))
        frames = g["frame"].reshape(1, H * 3 // 2, 2 * W)
        want = g["s8_correct"] if correct else g["s8"]
    else:
        rng = np.random.default_rng(H)
        frames = np.concatenate([_frames(H, W, 1, seed=60), rng.integers(0, 256, (1, H * 3 // 2, 2 * W), dtype=np.uint8)])
        want = _s8(frames, H, W, correct)
    # FLAG_NO_HEADFUSE: the pipeline variant that materialises the image tensor (the default path reads the s8 tensor directly)
    m = _model(H, W, K=3, D=4, max_batch=2, flags=capi.FLAG_KEEP_STAGES | capi.FLAG_NO_HEADFUSE | (capi.FLAG_CORRECT_CHROMA if correct else 0))
    got = m.pre_nv12_gpu(frames)
    assert got.dtype == np.int8 and got.shape == want.shape
    assert (got == want).all()                                          # byte for byte against the oracle / golden bytes
    left, right = capi.pre_split_nv12(frames[0].ravel(), H, 2 * W)
    assert (got[:1] == capi.pre_cvt_nv12_to_tensor(left, right, W, H, correct)).all()      # and against the host C function
    # the tensor the network actually reads: image stage = s8 / 128 exactly (left views then right views)
    m.infer_nv12(frames)
    img = m.debug_read("img")                                           # [2B, 3, Hp, Wp]
    B = frames.shape[0]
    ref = np.concatenate([want[:, :3], want[:, 3:]]).astype(np.float32) / 128.0
    assert (img[:, :, :H, :W] == ref).all()
    assert (img[:, :, H:, :] == 0).all() and (img[:, :, :, W:] == 0).all()
    # the s8 entry point writes the identical image
    q_old = m.infer(want)
    assert (m.debug_read("img")[:, :, :H, :W] == ref).all()
    m.close()
    # default path (no image tensor: firstconv.0 and the refinement heads read the s8 tensor): same bytes from the GPU pre-process,
    # and the camera-frame entry gives the same output as the tensor entry
    m = _model(H, W, K=3, D=4, max_batch=2, flags=capi.FLAG_CORRECT_CHROMA if correct else 0)
    assert (m.pre_nv12_gpu(frames) == want).all()
    q_new = m.infer(want)
    assert (m.infer_nv12(frames) == q_new).all()
    m.close()
    px = lambda q: q.astype(np.float64) * arch.OUT_SCALE * arch.OUT_NORM
    assert np.abs(px(q_new) - px(q_old)).mean() < 1e-3        # the two pipeline variants agree to the parity bar


def test_nv12_async_matches_sync_and_keeps_order(built_lib):
    H, W, N = 64, 96, 12
    frames = _frames(H, W, N, seed=900)
    s8 = _s8(frames, H, W)
    m = _model(H, W, max_batch=4, task_num=4)
    want = m.infer(s8)
    outs = [np.zeros((1, 1, H, W), np.int32) for _ in range(N)]
    order = []
    lock = threading.Lock()

    def cb(i):
        def f(status, stat):
            with lock:
                order.append((i, status))
        return f

    for i in range(N):                     # even calls: raw frames, odd calls: the s8 tensor - never merged into one pass
        if i % 2 == 0:
            m.infer_nv12_async(frames[i:i + 1], outs[i], cb(i))
        else:
            m.infer_async(s8[i:i + 1], outs[i], cb(i))
    m.wait_all()
    assert [i for i, _ in order] == list(range(N)) and all(s == 0 for _, s in order)
    for i in range(N):
        assert (outs[i] == want[i:i + 1]).all()
    # all NV12: the worker may merge them into passes of up to 4 pairs; results stay bit-identical
    p0 = m.pass_count()
    for i in range(N):
        m.infer_nv12_async(frames[i:i + 1], outs[i])
    m.wait_all()
    assert m.pass_count() - p0 <= N
    for i in range(N):
        assert (outs[i] == want[i:i + 1]).all()
    m.close()


def test_callback_may_resubmit_with_all_slots_taken(built_lib):
    """capi.cu retire_oldest frees the task slots before the callbacks run: a callback that blocks for a slot
    (timeout -1) while task_num calls are queued behind it must not deadlock the worker thread."""
    H, W = 48, 64
    frames = _frames(H, W, 2, seed=31)
    s8 = _s8(frames, H, W)
    m = _model(H, W, D=6, max_batch=1, task_num=2)
    want = m.infer(s8)
    outs = [np.zeros((1, 1, H, W), np.int32) for _ in range(8)]
    left = [6]
    done = threading.Event()

    def chain(status, stat):
        assert status == 0
        if left[0] > 0:
            left[0] -= 1
            k = 2 + (5 - left[0])
            m.infer_async(s8[k % 2:k % 2 + 1], outs[k], chain if left[0] > 0 else (lambda s, st: done.set()), timeout_ms=-1)

    m.infer_async(s8[0:1], outs[0], chain)
    m.infer_async(s8[1:2], outs[1], None)
    assert done.wait(timeout=60), "worker thread deadlocked on a resubmitting callback"
    m.wait_all()
    for k in range(8):
        assert (outs[k] == want[k % 2:k % 2 + 1]).all()
    m.close()


def test_set_weights_is_transactional_for_bad_blobs(built_lib):
    from hobot_stereonet_b200 import SnbError, capi
    H, W = 48, 64
    s8 = _s8(_frames(H, W, 1, seed=77), H, W)
    m = _model(H, W, D=6)
    q0 = m.infer(s8)
    good = weights.make_blob(3, seed=99)
    for bad in (good[:4096], weights.make_blob(4, seed=99), b"junk" * 100):
        with pytest.raises(SnbError) as e:
            m.set_weights(bad)
        assert e.value.code == capi.SNB_ERR_MODEL
        assert (m.infer(s8) == q0).all()          # the old model is still installed and still runs
    m.set_weights(good)
    q1 = m.infer(s8)
    assert (q1 != q0).any()
    m2 = _model(H, W, D=6, seed=99)
    assert (m2.infer(s8) == q1).all()
    m.close(); m2.close()


def test_caller_stream_pass_is_ordered_before_the_next_pass(built_lib):
    """snb_infer_device on a caller's stream returns without synchronising; every pass shares one scratch set, so the
    next pass (any entry point, the context's own stream) must wait for it on the device."""
    import torch
    H, W, B = 256, 512, 2
    frames = _frames(H, W, 2 * B, seed=5)
    s8 = _s8(frames, H, W)
    m = _model(H, W, D=12, max_batch=B)
    want = m.infer(s8)
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(s8).to(dev)
    d_out = torch.zeros((2 * B, 1, H, W), dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(dev)
    for rep in range(5):
        d_out.zero_()
        torch.cuda.synchronize()
        m.infer_device(d_in[:B], d_out[:B], B, st.cuda_stream)          # pass 1: caller's stream, not synchronised
        q2 = m.infer(s8[B:])                                              # pass 2: the context's stream, right behind it
        st.synchronize()
        assert (d_out[:B].cpu().numpy() == want[:B]).all() and (q2 == want[B:]).all(), rep
    m.close()


def test_pool_single_and_multi_gpu(built_lib):
    import torch
    from hobot_stereonet_b200 import Pool, capi
    H, W, N = 64, 96, 16
    ndev = torch.cuda.device_count()
    frames = _frames(H, W, N, seed=1000)
    s8 = _s8(frames, H, W)
    blob = weights.make_blob(3, seed=1234)
    m = _model(H, W, max_batch=4)
    want = m.infer(s8)
    m.close()
    for devs in ([0], list(range(ndev))):
        p = Pool(H, W, 3, 8, devices=devs, max_batch=4, weights=blob)
        assert p.size() == len(devs)
        assert (p.infer(s8) == want).all()                                # one call, contiguous shards over the replicas
        outs = [np.zeros((1, 1, H, W), np.int32) for _ in range(N)]
        for i in range(N):
            p.infer_async(frames[i:i + 1] if i % 2 else s8[i:i + 1], outs[i], nv12=bool(i % 2))
        p.wait_all()
        for i in range(N):
            assert (outs[i] == want[i:i + 1]).all()
        st = p.stat()
        assert st["n_devices"] == len(devs) and st["weight_bytes"] == len(blob) and sum(st["calls"]) == N + len(devs)
        if len(devs) > 1:
            assert st["broadcast_ms"] > 0 and min(st["calls"]) >= 1       # every replica got the weights and served calls
        p.close()
        if ndev == 1:
            break
