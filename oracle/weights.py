"""TEST INFRASTRUCTURE — seeded synthetic weights + the weight-blob container.

The reference's float weights are unrecoverable (int8 "kcp" BPU layout,
SURVEY.md §2.3), so both the oracle and the CUDA path consume the same seeded
blob.  `model_file` (stereonet_node.cpp:131-136) names such a blob in this
build instead of an .hbm.

Blob layout (little endian):
  0   char[8]  magic "SNB2WGT1"
  8   u32 version(1), u32 K, u32 n_tensors, u32 reserved
  24  n_tensors x { char name[64]; u32 ndim; u32 dims[5]; u64 offset; u64 nbytes }   (104 B each)
  ..  pad to 64 B, then fp32 data; `offset` is relative to the start of the data section
Conv weights are stored in the canonical [Cout, Cin, (kd,) kh, kw] order.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

from .arch import conv_specs

MAGIC = b"SNB2WGT1"
ENTRY = struct.Struct("<64sI5IQQ")
HEADER = struct.Struct("<8sIIII")


def generate(K: int, seed: int = 1234) -> Dict[str, np.ndarray]:
    """He-normal weights (std = gain*sqrt(2/fan_in)), small random biases."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for spec in conv_specs(K):
        kk = int(np.prod(spec.k))
        fan_in = spec.cin * kk
        std = spec.gain * np.sqrt(2.0 / fan_in)
        w = rng.standard_normal((spec.cout, spec.cin) + tuple(spec.k)).astype(np.float32) * np.float32(std)
        b = rng.standard_normal(spec.cout).astype(np.float32) * np.float32(0.05 * min(spec.gain, 1.0))
        out[spec.name + ".weight"] = w
        out[spec.name + ".bias"] = b
    return out


def to_blob(tensors: Dict[str, np.ndarray], K: int) -> bytes:
    names = list(tensors.keys())
    table = bytearray()
    data = bytearray()
    for n in names:
        a = np.ascontiguousarray(tensors[n], dtype="<f4")
        assert a.ndim <= 5 and len(n.encode()) < 64
        dims = list(a.shape) + [1] * (5 - a.ndim)
        off = len(data)
        data += a.tobytes()
        data += b"\0" * (-len(data) % 64)
        table += ENTRY.pack(n.encode(), a.ndim, *dims, off, a.nbytes)
    head = HEADER.pack(MAGIC, 1, K, len(names), 0) + bytes(table)
    head += b"\0" * (-len(head) % 64)
    return bytes(head) + bytes(data)


def from_blob(blob: bytes):
    magic, ver, K, n, _ = HEADER.unpack_from(blob, 0)
    if magic != MAGIC or ver != 1:
        raise ValueError("not a SNB2WGT1 blob")
    pos = HEADER.size
    entries = []
    for _ in range(n):
        name, ndim, d0, d1, d2, d3, d4, off, nbytes = ENTRY.unpack_from(blob, pos)
        pos += ENTRY.size
        entries.append((name.rstrip(b"\0").decode(), (d0, d1, d2, d3, d4)[:ndim], off, nbytes))
    base = (pos + 63) // 64 * 64
    out = {}
    for name, shape, off, nbytes in entries:
        out[name] = np.frombuffer(blob, dtype="<f4", count=nbytes // 4, offset=base + off).reshape(shape).copy()
    return K, out


def make_blob(K: int, seed: int = 1234) -> bytes:
    return to_blob(generate(K, seed), K)
