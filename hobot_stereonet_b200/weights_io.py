"""Weight blobs for `model_file` (import path for real float weights).

The reference ships its network only as a compiled BPU binary; its float model is the HAT StereoNet of Horizon
OpenExplorer (reference README.md:5), a PyTorch model whose convolutions are followed by BatchNorm.  This module turns
such a checkpoint into the SNB2WGT1 blob `snb_create` loads (`snb_config.model_file` / `.weights`, the replacement of
`dnn_node_para_ptr_->model_file`, stereonet_node.cpp:131-136):

  * `fold_bn`            conv + BatchNorm (eval mode) -> one conv with bias, in float64, rounded once to fp32;
  * `import_state_dict`  walks the layer list the library itself expects (the tensor table of a synthesized blob),
                         finds each layer's convolution and BatchNorm in the checkpoint (explicit map or naming
                         heuristics), folds, checks shapes, and reports what it could not find;
  * `write_blob` / `read_blob`   the container (layout below; parsed by csrc/net.cu).

Blob layout (little endian): char[8] "SNB2WGT1"; u32 version = 1, u32 K, u32 n_tensors, u32 reserved; n_tensors x
{char name[64]; u32 ndim; u32 dims[5]; u64 offset; u64 nbytes}; pad to 64 B; fp32 data (offsets relative to its start,
64-byte aligned).  Conv weights are [Cout, Cin, (kd,) kh, kw].  No GPU and no torch needed: tensors are numpy arrays
(`torch.load(...)` values are converted by the CLI, tools/import_weights.py).
"""
from __future__ import annotations

import struct
from typing import Dict, Iterable, List, Mapping, Optional, Tuple

import numpy as np

MAGIC = b"SNB2WGT1"
_HEADER = struct.Struct("<8sIIII")
_ENTRY = struct.Struct("<64sI5IQQ")


def write_blob(tensors: Mapping[str, np.ndarray], K: int) -> bytes:
    table, data = bytearray(), bytearray()
    for name, t in tensors.items():
        a = np.ascontiguousarray(t, dtype="<f4")
        if a.ndim > 5 or len(name.encode()) >= 64:
            raise ValueError(f"tensor {name}: at most 5 dims and 63 name bytes")
        dims = list(a.shape) + [1] * (5 - a.ndim)
        off = len(data)
        data += a.tobytes()
        data += b"\0" * (-len(data) % 64)
        table += _ENTRY.pack(name.encode(), a.ndim, *dims, off, a.nbytes)
    head = _HEADER.pack(MAGIC, 1, K, len(tensors), 0) + bytes(table)
    head += b"\0" * (-len(head) % 64)
    return bytes(head) + bytes(data)


def read_blob(blob: bytes) -> Tuple[int, Dict[str, np.ndarray]]:
    magic, ver, K, n, _ = _HEADER.unpack_from(blob, 0)
    if magic != MAGIC or ver != 1:
        raise ValueError("not a SNB2WGT1 blob")
    pos = _HEADER.size
    entries = []
    for _ in range(n):
        name, ndim, d0, d1, d2, d3, d4, off, nbytes = _ENTRY.unpack_from(blob, pos)
        pos += _ENTRY.size
        entries.append((name.rstrip(b"\0").decode(), (d0, d1, d2, d3, d4)[:ndim], off, nbytes))
    base = (pos + 63) // 64 * 64
    out = {}
    for name, shape, off, nbytes in entries:
        if base + off + nbytes > len(blob):
            raise ValueError(f"tensor {name} out of range")
        out[name] = np.frombuffer(blob, dtype="<f4", count=nbytes // 4, offset=base + off).reshape(shape).copy()
    return K, out


def expected_layers(K: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """(conv name, weight shape) of every convolution the library expects for K refinement stages, in execution order -
    read from the tensor table of the library's own synthesized blob, so this list cannot drift from csrc/."""
    from . import capi
    _, t = read_blob(capi.synthesize_weights(K))
    return [(n[:-len(".weight")], tuple(a.shape)) for n, a in t.items() if n.endswith(".weight")]


def fold_bn(w: np.ndarray, b: Optional[np.ndarray], gamma: np.ndarray, beta: np.ndarray, mean: np.ndarray, var: np.ndarray,
            eps: float = 1e-5) -> Tuple[np.ndarray, np.ndarray]:
    """y = BN(conv(x, w) + b) in eval mode  ==  conv(x, w') + b'  with
    w' = w * gamma / sqrt(var + eps) per output channel, b' = (b - mean) * gamma / sqrt(var + eps) + beta."""
    w64 = np.asarray(w, np.float64)
    s = np.asarray(gamma, np.float64) / np.sqrt(np.asarray(var, np.float64) + eps)
    b64 = np.zeros(w64.shape[0]) if b is None else np.asarray(b, np.float64)
    wf = w64 * s.reshape((-1,) + (1,) * (w64.ndim - 1))
    bf = (b64 - np.asarray(mean, np.float64)) * s + np.asarray(beta, np.float64)
    return wf.astype(np.float32), bf.astype(np.float32)


_CONV_SUFFIXES = ("", ".conv", ".0", ".conv.0")
_BN_SUFFIXES = (".bn", ".1", ".norm", ".conv.1", "_bn")


def _find(sd: Mapping[str, np.ndarray], base: str, suffixes: Iterable[str], key: str) -> Optional[str]:
    for s in suffixes:
        if base + s + key in sd:
            return base + s
    return None


def import_state_dict(sd: Mapping[str, np.ndarray], K: int, name_map: Optional[Mapping[str, Mapping[str, str]]] = None,
                      eps: float = 1e-5, layers: Optional[List[Tuple[str, Tuple[int, ...]]]] = None):
    """Checkpoint (name -> array) -> ({layer.weight/.bias -> fp32 array}, report).

    name_map[layer] = {"conv": <checkpoint prefix of the convolution>, "bn": <prefix of its BatchNorm, optional>} overrides
    the heuristics, which try `<layer>`, `<layer>.conv`, `<layer>.0` for the convolution and `<layer>.bn`, `<layer>.1`,
    `<layer>.norm`, `<layer>_bn` for the BatchNorm (`weight`, `bias`, `running_mean`, `running_var`).  A layer without
    BatchNorm keeps its own bias (zero if it has none).  Raises on a missing layer or a shape mismatch: a silently
    half-imported network is worse than none."""
    layers = layers if layers is not None else expected_layers(K)
    out: Dict[str, np.ndarray] = {}
    report = {"folded": [], "plain": [], "unused": []}
    used = set()
    for layer, shape in layers:
        m = (name_map or {}).get(layer, {})
        conv = m.get("conv") or _find(sd, layer, _CONV_SUFFIXES, ".weight")
        if conv is None or conv + ".weight" not in sd:
            raise KeyError(f"no convolution found for layer {layer} (tried {[layer + s for s in _CONV_SUFFIXES]}; pass name_map)")
        w = np.asarray(sd[conv + ".weight"], np.float32)
        if tuple(w.shape) != tuple(shape):
            raise ValueError(f"layer {layer}: checkpoint tensor {conv}.weight has shape {tuple(w.shape)}, the network needs {tuple(shape)}")
        b = np.asarray(sd[conv + ".bias"], np.float32) if conv + ".bias" in sd else None
        used.update({conv + ".weight", conv + ".bias"})
        bn = m.get("bn")
        if bn is None and "bn" not in m:
            cands = [layer + s for s in _BN_SUFFIXES] + ([conv[:-len(".conv")] + ".bn"] if conv.endswith(".conv") else [])
            bn = next((c for c in cands if c != conv and c + ".running_var" in sd), None)
        if bn:
            g, be = sd[bn + ".weight"], sd[bn + ".bias"]
            w, b = fold_bn(w, b, g, be, sd[bn + ".running_mean"], sd[bn + ".running_var"], eps)
            used.update({bn + k for k in (".weight", ".bias", ".running_mean", ".running_var", ".num_batches_tracked")})
            report["folded"].append(layer)
        else:
            b = np.zeros(shape[0], np.float32) if b is None else b
            report["plain"].append(layer)
        out[layer + ".weight"], out[layer + ".bias"] = w, b
    report["unused"] = sorted(k for k in sd if k not in used)
    return out, report
