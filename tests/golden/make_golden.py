"""Generates the committed fixtures under tests/golden/ (run from the repo root:
`python tests/golden/make_golden.py`).  Nothing here reads /root/reference: the reference ships no
input/output tensors for this path, so the fixtures are (a) a literal loop-level
transcription of its pre-process applied to a seeded frame and (b) outputs of the fp32 oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import arch, prepost_ref as pp, synth, weights  # noqa: E402
from oracle.stereonet_ref import Oracle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

NET_CASES = {           # name: (H, W, K, D, batch)
    "net_64x96_k3_d8": (64, 96, 3, 8, 1),
    "net_50x70_k2_d6": (50, 70, 2, 6, 2),     # ragged: padded to 52x72
    "net_64x128_k4_d4": (64, 128, 4, 4, 1),   # the deployed model's K
    "net_40x48_k3_d12": (40, 48, 3, 12, 1),   # D > cost-volume width (6): all-zero slices
}


def literal_pre(frame, h, w):
    """Loop-level transcription of stereonet_node.cpp:702-738 + preprocess.h:128-155 +
    preprocess.cpp:1032-1040 with the float Quantize of :1131-1136."""
    rows = frame.reshape(h * 3 // 2, 2 * w)
    views = [rows[:, :w].ravel(), rows[:, w:].ravel()]
    planes = []
    for buf in views:
        o = np.zeros((3, h, w), np.uint8)
        o[0] = buf[: w * h].reshape(h, w)
        for i in range(0, h, 2):
            for j in range(0, w, 2):
                o[1, i:i + 2, j:j + 2] = buf[w * h + i // 2 * w // 2 + j // 2]
                o[2, i:i + 2, j:j + 2] = buf[w * h + w * h // 4 + i // 2 * w // 2 + j // 2]
        planes.append(o)
    x = np.concatenate(planes).astype(np.float32)
    v = np.floor(((x - np.float32(128.0)) / np.float32(128.0)) / np.float32(0.0078125) + np.float32(0.5))
    return np.clip(v, -128, 127).astype(np.int8)[None]


def main():
    h, w = 16, 24
    frame = synth.frame(h, w, 16, seed=11)
    l, r = pp.split_side_by_side_nv12(frame, h, 2 * w)
    np.savez_compressed(os.path.join(OUT, "prepost_16x24.npz"), frame=frame, s8=literal_pre(frame, h, w),
                        s8_correct=pp.cvt_nv12_to_tensor(l, r, w, h, correct_chroma=True))
    for name, (H, W, K, D, B) in NET_CASES.items():
        cfg = arch.Config(H, W, K, D)
        wts = weights.generate(K, seed=1234)
        frames = np.stack([synth.frame(H, W, cfg.max_disp, seed=100 + i) for i in range(B)])
        s8 = np.concatenate([pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(f, H, 2 * W), W, H) for f in frames])
        o = Oracle(cfg, wts)
        dump = {}
        dn = o.forward_norm(s8, dump).numpy()
        q = o.forward_s32(s8)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), cfg=np.array([H, W, K, D, B]), frames=frames, s8=s8,
                            q=q, dn=dn.astype(np.float32), disp0=dump["disp0"].numpy(),
                            cost=dump["cost"].numpy(), cat=dump["cat"].numpy())
        print(name, "disp px mean", float(dn.mean() * cfg.max_disp))


if __name__ == "__main__":
    main()
