"""checkpoint import (BatchNorm folding + SNB2WGT1 container), CPU only.
The folding is checked against PyTorch's own conv -> batch_norm (eval) on random data; the container against the oracle's
independent reader and against the blob the C library synthesizes; the importer on a synthetic conv+BN checkpoint of the
whole network whose folded result is then run through the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from hobot_stereonet_b200 import weights_io
from oracle import arch, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fold_bn_matches_torch_conv_then_batchnorm():
    g = torch.Generator().manual_seed(5)
    for shape in [(8, 3, 3, 3), (4, 6, 3, 3, 3), (5, 7, 1, 1)]:
        w = torch.randn(shape, generator=g) * 0.3
        b = torch.randn(shape[0], generator=g) * 0.1
        gamma, beta = torch.rand(shape[0], generator=g) + 0.5, torch.randn(shape[0], generator=g) * 0.2
        mean, var = torch.randn(shape[0], generator=g) * 0.3, torch.rand(shape[0], generator=g) + 0.1
        x = torch.randn((2, shape[1]) + (6,) * (len(shape) - 2), generator=g)
        conv = F.conv3d if len(shape) == 5 else F.conv2d
        ref = F.batch_norm(conv(x, w, b, padding=shape[-1] // 2), mean, var, gamma, beta, training=False, eps=1e-5)
        wf, bf = weights_io.fold_bn(w.numpy(), b.numpy(), gamma.numpy(), beta.numpy(), mean.numpy(), var.numpy(), 1e-5)
        out = conv(x, torch.from_numpy(wf), torch.from_numpy(bf), padding=shape[-1] // 2)
        assert torch.allclose(out, ref, atol=2e-6, rtol=1e-5)
        # no conv bias
        wf2, bf2 = weights_io.fold_bn(w.numpy(), None, gamma.numpy(), beta.numpy(), mean.numpy(), var.numpy(), 1e-5)
        ref2 = F.batch_norm(conv(x, w, None, padding=shape[-1] // 2), mean, var, gamma, beta, training=False, eps=1e-5)
        assert torch.allclose(conv(x, torch.from_numpy(wf2), torch.from_numpy(bf2), padding=shape[-1] // 2), ref2, atol=2e-6, rtol=1e-5)


def test_blob_container_agrees_with_oracle_and_library(built_lib):
    from hobot_stereonet_b200 import capi
    t = weights.generate(3, seed=77)
    blob = weights_io.write_blob(t, 3)
    assert blob == weights.to_blob(t, 3)                              # two independent writers, same bytes
    K, back = weights.from_blob(blob)
    assert K == 3 and all((back[k] == t[k]).all() for k in t)
    K2, lib_t = weights_io.read_blob(capi.synthesize_weights(4))      # the C library's writer (csrc/weights_host.cpp)
    assert K2 == 4
    specs = arch.conv_specs(4)
    assert [n for n, _ in weights_io.expected_layers(4)] == [s.name for s in specs]
    for s in specs:
        assert lib_t[s.name + ".weight"].shape == (s.cout, s.cin) + tuple(s.k)
    with pytest.raises(ValueError):
        weights_io.read_blob(b"NOTABLOB" + blob[8:])


def _synthetic_checkpoint(K, seed, style, gains=False):
    """conv (no bias) + BatchNorm for every layer, named the way `style` says.  gains: scale every layer like the seeded blob
    (oracle/arch.py ConvSpec.gain), so that activations stay O(1) through the 25 residual blocks as in a trained network."""
    rng = np.random.default_rng(seed)
    folded_ref = {}
    sd = {}
    for s in arch.conv_specs(K):
        w = (rng.standard_normal((s.cout, s.cin) + tuple(s.k)) * np.sqrt(2.0 / (s.cin * np.prod(s.k))) * (s.gain if gains else 1.0)).astype(np.float32)
        gamma, beta = (rng.random(s.cout) + 0.5).astype(np.float32), (rng.standard_normal(s.cout) * 0.05).astype(np.float32)
        mean, var = (rng.standard_normal(s.cout) * 0.1).astype(np.float32), (rng.random(s.cout) + 0.5).astype(np.float32)
        conv, bn = {"seq": (s.name + ".0", s.name + ".1"), "named": (s.name + ".conv", s.name + ".bn")}[style]
        sd[conv + ".weight"] = w
        sd.update({bn + ".weight": gamma, bn + ".bias": beta, bn + ".running_mean": mean, bn + ".running_var": var,
                   bn + ".num_batches_tracked": np.array(100)})
        folded_ref[s.name] = weights_io.fold_bn(w, None, gamma, beta, mean, var)
    return sd, folded_ref


@pytest.mark.parametrize("style", ["seq", "named"])
def test_import_state_dict_folds_every_layer(style):
    K = 3
    layers = [(s.name, (s.cout, s.cin) + tuple(s.k)) for s in arch.conv_specs(K)]
    sd, ref = _synthetic_checkpoint(K, 3, style)
    sd["extra.head.weight"] = np.zeros(3, np.float32)
    tensors, report = weights_io.import_state_dict(sd, K, layers=layers)
    assert len(report["folded"]) == len(layers) and not report["plain"] and report["unused"] == ["extra.head.weight"]
    for name, (w, b) in ref.items():
        assert (tensors[name + ".weight"] == w).all() and (tensors[name + ".bias"] == b).all()
    # the imported blob drives the oracle like any other blob
    from oracle.stereonet_ref import Oracle
    from oracle import synth, prepost_ref as pp
    cfg = arch.Config(32, 64, K, 4)
    _, got = weights.from_blob(weights_io.write_blob(tensors, K))
    frame = synth.frame(cfg.H, cfg.W, cfg.max_disp, seed=1)
    s8 = pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(frame, cfg.H, 2 * cfg.W), cfg.W, cfg.H)
    disp = Oracle(cfg, got).forward_px(s8)
    assert disp.shape[-2:] == (cfg.H, cfg.W) and np.isfinite(disp).all()
    # wrong shape and missing layer are errors, not silent skips
    bad = dict(sd)
    k0 = next(k for k in bad if k.endswith(".weight") and bad[k].ndim == 4)
    bad[k0] = bad[k0][:, :, :1]
    with pytest.raises(ValueError):
        weights_io.import_state_dict(bad, K, layers=layers)
    missing = {k: v for k, v in sd.items() if not k.startswith("head.conv3d_alone")}
    with pytest.raises(KeyError):
        weights_io.import_state_dict(missing, K, layers=layers)


def test_import_cli_with_explicit_map(built_lib, tmp_path):
    K = 3
    sd, ref = _synthetic_checkpoint(K, 9, "named")
    # rename one layer to something the heuristics cannot guess and map it explicitly
    renamed = {}
    for k, v in sd.items():
        k = k.replace("backbone.firstconv.0.conv", "stem.c1").replace("backbone.firstconv.0.bn", "stem.n1")
        renamed["module." + k] = torch.from_numpy(np.asarray(v))
    ck = tmp_path / "ck.pth"
    torch.save({"state_dict": renamed}, ck)
    mp = tmp_path / "map.json"
    mp.write_text('{"backbone.firstconv.0": {"conv": "stem.c1", "bn": "stem.n1"}}')
    out = tmp_path / "model.snb"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "import_weights.py"), "--checkpoint", str(ck), "--K", str(K),
                        "--out", str(out), "--map", str(mp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    K2, t = weights_io.read_blob(out.read_bytes())
    assert K2 == K
    for name, (w, b) in ref.items():
        assert (t[name + ".weight"] == w).all() and (t[name + ".bias"] == b).all()


def test_import_cli_lists_expected_layers(built_lib):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "import_weights.py"), "--list", "--K", "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    specs = arch.conv_specs(4)
    assert len(lines) == len(specs)
    assert lines[0] == "backbone.firstconv.0 32x3x3x3" and lines[-1].startswith("head.refine.3.conv_out 1x32x3x3")


@pytest.mark.gpu
def test_imported_checkpoint_runs_on_gpu_like_the_oracle(built_lib, tmp_path):
    """SURVEY.md §8f rank 4 on the GPU: a conv + BatchNorm checkpoint goes through tools/import_weights.py, the resulting
    model_file drives both the CPU oracle and the tensor-core path (weights of another distribution than the seeded He-normal
    blob: per-channel BatchNorm scales, non-zero folded biases), and the two agree to the parity bar."""
    from hobot_stereonet_b200 import Model, capi
    from oracle import prepost_ref as pp, synth
    from oracle.stereonet_ref import Oracle
    K, H, W, D = 3, 96, 160, 8
    sd, _ = _synthetic_checkpoint(K, 21, "named", gains=True)
    ck = tmp_path / "ck.pth"
    torch.save({"state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}}, ck)
    out = tmp_path / "model.snb"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "import_weights.py"), "--checkpoint", str(ck), "--K", str(K), "--out", str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    blob = out.read_bytes()
    capi.weights_validate(blob, K)
    cfg = arch.Config(H, W, K, D)
    frame = synth.frame(H, W, cfg.max_disp, seed=4)
    s8 = pp.cvt_nv12_to_tensor_fast(*pp.split_side_by_side_nv12(frame, H, 2 * W), W, H)
    ref = Oracle(cfg, weights.from_blob(blob)[1]).forward_px(s8)
    m = Model(H, W, K, D, model_file=str(out), precision=capi.PREC_TC_F16X2)
    q = m.infer(s8)
    m.close()
    err = np.abs(q[:, 0].astype(np.float64) * arch.OUT_SCALE * arch.OUT_NORM - ref)
    print(f"imported checkpoint: mean EPE {err.mean():.3e} px, max {err.max():.3e} px; disparity range {ref.min():.2f}..{ref.max():.2f} px")
    assert np.isfinite(ref).all() and err.mean() <= 1e-3 and err.max() <= 2e-2
