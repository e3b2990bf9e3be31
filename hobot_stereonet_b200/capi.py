"""ctypes binding of libsnb200.so (C ABI: include/snb200.h).

There is no Python/CPU fallback: importing this module without the built library raises, and
creating a context without an sm_100 GPU returns SNB_ERR_CUDA from the library itself.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsnb200.so")

SNB_OK, SNB_ERR_INVALID, SNB_ERR_MODEL, SNB_ERR_CUDA, SNB_ERR_NOMEM, SNB_ERR_BUSY = 0, -1, -2, -3, -4, -5
PREC_FP32, PREC_TC_F16X2 = 0, 1
FLAG_DEFER_WEIGHTS = 1024
FLAG_NO_HEADFUSE = 64
FLAG_KEEP_STAGES, FLAG_NO_GRAPH, FLAG_CORRECT_CHROMA, FLAG_NO_TENSOR, FLAG_NO_FUSE, FLAG_NO_STREAM, FLAG_NO_HBMCONV, FLAG_NO_COALESCE = 1, 2, 4, 8, 16, 32, 128, 256
LAYOUT_NCHW, TENSOR_S8, TENSOR_S32 = 2, 1, 3


class SnbConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("K", C.c_int32),
                ("D", C.c_int32), ("max_batch", C.c_int32), ("device", C.c_int32), ("task_num", C.c_int32),
                ("precision", C.c_int32), ("flags", C.c_int32), ("model_file", C.c_char_p),
                ("weights", C.c_void_p), ("weights_bytes", C.c_uint64)]


class SnbTensorProps(C.Structure):
    _fields_ = [("valid_shape", C.c_int32 * 4), ("aligned_shape", C.c_int32 * 4), ("tensor_layout", C.c_int32),
                ("tensor_type", C.c_int32), ("scale_len", C.c_int32), ("scale", C.c_float), ("mem_size", C.c_uint64)]


class SnbRtStat(C.Structure):
    _fields_ = [("input_fps", C.c_float), ("output_fps", C.c_float), ("infer_time_ms", C.c_int32),
                ("fps_updated", C.c_int32), ("gpu_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
                ("kernel_launches", C.c_int32)]


class SnbPoolStat(C.Structure):
    _fields_ = [("n_devices", C.c_int32), ("device", C.c_int32 * 16), ("calls", C.c_int64 * 16), ("weight_bytes", C.c_uint64),
                ("broadcast_ms", C.c_float)]


class SnbKernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("ms", C.c_float), ("flops", C.c_double), ("bytes", C.c_double)]


DONE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(SnbRtStat))


def load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no fallback path)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u64, i64 = C.c_void_p, C.c_int32, C.c_uint64, C.c_int64
    sig = {
        "snb_create": (C.c_int, [C.POINTER(vp), C.POINTER(SnbConfig)]),
        "snb_destroy": (None, [vp]),
        "snb_set_weights": (C.c_int, [vp, vp, u64, C.c_int]),
        "snb_sys_alloc": (C.c_int, [C.POINTER(vp), u64]),
        "snb_sys_free": (None, [vp]),
        "snb_get_io": (C.c_int, [vp, C.POINTER(SnbTensorProps), C.POINTER(SnbTensorProps)]),
        "snb_get_model_input_size": (C.c_int, [vp, i32, C.POINTER(i32), C.POINTER(i32)]),
        "snb_infer": (C.c_int, [vp, vp, vp, i32]),
        "snb_infer_async": (C.c_int, [vp, vp, vp, i32, DONE_FN, vp, i32]),
        "snb_wait_all": (C.c_int, [vp]),
        "snb_infer_device": (C.c_int, [vp, vp, vp, i32, vp]),
        "snb_infer_nv12": (C.c_int, [vp, vp, vp, i32]),
        "snb_get_rt_stat": (C.c_int, [vp, C.POINTER(SnbRtStat)]),
        "snb_get_pass_count": (i64, [vp]),
        "snb_last_error": (C.c_char_p, [vp]),
        "snb_version": (C.c_char_p, []),
        "snb_debug_read": (i64, [vp, C.c_char_p, vp, u64, C.POINTER(i32 * 5)]),
        "snb_profile_pass": (C.c_int, [vp, i32, C.POINTER(SnbKernelTime), i32]),
        "snb_pre_split_nv12": (C.c_int, [vp, i32, i32, vp, vp]),
        "snb_pre_yuv420_to_yuv444": (C.c_int, [vp, vp, i32, i32, i32]),
        "snb_pre_cvt_nv12_to_tensor": (C.c_int, [vp, vp, i32, i32, i32, vp]),
        "snb_pre_quantize": (C.c_int8, [C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]),
        "snb_post_pack": (i64, [vp, u64, vp, u64, vp, u64]),
        "snb_post_parse_depth": (C.c_int, [vp, i64, C.c_float, vp]),
        "snb_weights_synthesize": (i64, [i32, u64, vp, u64]),
        "snb_post_depth_color": (C.c_int, [vp, vp, i32, C.c_float, vp, vp, i32]),
        "snb_infer_nv12_async": (C.c_int, [vp, vp, vp, i32, DONE_FN, vp, i32]),
        "snb_pre_nv12_gpu": (C.c_int, [vp, vp, i32, vp]),
        "snb_pre_nv12_to_bgr": (C.c_int, [vp, i32, i32, vp]),
        "snb_jpeg_encode_nv12": (i64, [vp, i32, i32, i32, vp, u64]),
        "snb_weights_validate": (C.c_int, [vp, u64, i32]),
        "snb_shard_range": (C.c_int, [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]),
        "snb_pool_create": (C.c_int, [C.POINTER(vp), C.POINTER(SnbConfig), C.POINTER(i32), i32]),
        "snb_pool_destroy": (None, [vp]),
        "snb_pool_size": (i32, [vp]),
        "snb_pool_ctx": (vp, [vp, i32]),
        "snb_pool_infer_async": (C.c_int, [vp, vp, vp, i32, DONE_FN, vp, i32]),
        "snb_pool_infer_nv12_async": (C.c_int, [vp, vp, vp, i32, DONE_FN, vp, i32]),
        "snb_pool_infer": (C.c_int, [vp, vp, vp, i32]),
        "snb_pool_wait_all": (C.c_int, [vp]),
        "snb_pool_get_stat": (C.c_int, [vp, C.POINTER(SnbPoolStat)]),
        "snb_pool_last_error": (C.c_char_p, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = load()
    return _lib


class SnbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"snb200 error {code}: {msg}")
        self.code = code


def _ptr(a) -> int:
    """Address of a numpy array or torch tensor (host or device)."""
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return a.data_ptr()
    return int(a)


class Model:
    """One model instance on one GPU — what DnnNode::Init()/GetModel() hand the reference node."""

    def __init__(self, height: int, width: int, K: int, D: int, *, max_batch: int = 1, device: int = 0,
                 task_num: int = 4, precision: int = PREC_FP32, flags: int = 0,
                 model_file: Optional[str] = None, weights: Optional[bytes] = None):
        self._l = lib()
        self._h = C.c_void_p()
        self._cbs = {}
        self._pending = []
        cfg = SnbConfig(C.sizeof(SnbConfig), height, width, K, D, max_batch, device, task_num, precision, flags,
                        model_file.encode() if model_file else None, None, 0)
        self._wbuf = None
        if weights is not None:
            self._wbuf = C.create_string_buffer(weights, len(weights))
            cfg.weights = C.cast(self._wbuf, C.c_void_p)
            cfg.weights_bytes = len(weights)
        r = self._l.snb_create(C.byref(self._h), C.byref(cfg))
        if r != SNB_OK:
            raise SnbError(r, (self._l.snb_last_error(None) or b"").decode())
        self.H, self.W, self.K, self.D, self.max_batch = height, width, K, D, max_batch

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._l.snb_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _check(self, r: int):
        if r < 0:
            raise SnbError(r, (self._l.snb_last_error(self._h) or b"").decode())
        return r

    def io_props(self):
        i, o = SnbTensorProps(), SnbTensorProps()
        self._check(self._l.snb_get_io(self._h, C.byref(i), C.byref(o)))
        return i, o

    def model_input_size(self):
        w, h = C.c_int32(), C.c_int32()
        self._check(self._l.snb_get_model_input_size(self._h, 0, C.byref(w), C.byref(h)))
        return w.value, h.value

    def set_weights(self, blob, nbytes: Optional[int] = None, is_device: bool = False):
        if isinstance(blob, (bytes, bytearray)):
            buf = C.create_string_buffer(bytes(blob), len(blob))
            self._check(self._l.snb_set_weights(self._h, C.cast(buf, C.c_void_p), len(blob), 0))
        else:
            self._check(self._l.snb_set_weights(self._h, _ptr(blob), nbytes, 1 if is_device else 0))

    def infer(self, s8, out=None):
        """s8: int8 [B,6,H,W] host array/tensor -> int32 [B,1,H,W] (DnnNode::Run, sync)."""
        B = s8.shape[0]
        if out is None:
            out = np.empty((B, 1, self.H, self.W), np.int32)
        self._check(self._l.snb_infer(self._h, _ptr(s8), _ptr(out), B))
        return out

    def infer_nv12(self, frames, out=None):
        """frames: uint8 [B, H*3/2, 2W] side-by-side NV12 camera frames (host)."""
        B = frames.shape[0]
        if out is None:
            out = np.empty((B, 1, self.H, self.W), np.int32)
        self._check(self._l.snb_infer_nv12(self._h, _ptr(frames), _ptr(out), B))
        return out

    def infer_device(self, d_in, d_out, batch: int, stream: int = 0):
        self._check(self._l.snb_infer_device(self._h, _ptr(d_in), _ptr(d_out), batch, stream or None))

    def pre_nv12_gpu(self, frames) -> np.ndarray:
        """CvtNV12Data2Tensors on the GPU: uint8 [B, H*3/2, 2W] frames -> the int8 [B,6,H,W] tensor (host)."""
        B = frames.shape[0]
        out = np.empty((B, 6, self.H, self.W), np.int8)
        self._check(self._l.snb_pre_nv12_gpu(self._h, _ptr(frames), B, _ptr(out)))
        return out

    def infer_async(self, s8, out, done=None, timeout_ms: int = -1, nv12: bool = False):
        fn = self._l.snb_infer_nv12_async if nv12 else self._l.snb_infer_async
        if done is None:
            # no Python callback: the library's worker thread never has to take the GIL for this call.  The buffers
            # stay referenced until wait_all() (the caller must not free them earlier anyway).
            self._pending.append((s8, out))
            if len(self._pending) > 4096:
                del self._pending[:2048]             # far more than task_num calls ago: long finished
            self._check(fn(self._h, _ptr(s8), _ptr(out), s8.shape[0], C.cast(None, DONE_FN), None, timeout_ms))
            return
        key = id(out)

        def _cb(user, status, stat):
            st = stat.contents
            try:
                done(status, {"gpu_ms": st.gpu_ms, "infer_time_ms": st.infer_time_ms})
            finally:
                self._cbs.pop(key, None)

        cb = DONE_FN(_cb)
        self._cbs[key] = (cb, s8, out)      # keep buffers and the thunk alive until the callback fired
        r = fn(self._h, _ptr(s8), _ptr(out), s8.shape[0], cb, None, timeout_ms)
        if r < 0:
            self._cbs.pop(key, None)
            self._check(r)

    def infer_nv12_async(self, frames, out, done=None, timeout_ms: int = -1):
        """frames: uint8 [B, H*3/2, 2W] raw camera frames (host, ideally pinned); the pre-process runs inside the pass."""
        self.infer_async(frames, out, done, timeout_ms, nv12=True)

    def wait_all(self):
        self._check(self._l.snb_wait_all(self._h))
        self._pending.clear()

    def pass_count(self) -> int:
        return int(self._l.snb_get_pass_count(self._h))

    def rt_stat(self) -> SnbRtStat:
        s = SnbRtStat()
        self._check(self._l.snb_get_rt_stat(self._h, C.byref(s)))
        return s

    def debug_read(self, name: str) -> np.ndarray:
        shape = (C.c_int32 * 5)()
        n = self._check(self._l.snb_debug_read(self._h, name.encode(), None, 0, C.byref(shape)))
        dst = np.empty(n, np.float32)
        self._check(self._l.snb_debug_read(self._h, name.encode(), dst.ctypes.data, n, C.byref(shape)))
        N, Cc, D, H, W = list(shape)
        a = dst.reshape(N, Cc, D, H, W)
        return a[:, :, 0] if D == 1 else a

    def depth_color(self, q: np.ndarray, alpha: float = 11.0):
        """ParseTensor on the GPU (parser.cpp:79-118): q int32 [B,1,H,W] host -> (depth f32 [B,H,W], bgr u8 [B,H,W,3])."""
        B = q.shape[0]
        depth = np.empty((B, self.H, self.W), np.float32)
        bgr = np.empty((B, self.H, self.W, 3), np.uint8)
        self._check(self._l.snb_post_depth_color(self._h, _ptr(np.ascontiguousarray(q, np.int32)), B, alpha, _ptr(depth), _ptr(bgr), 0))
        return depth, bgr

    def profile_pass(self, batch: int = 1):
        arr = (SnbKernelTime * 512)()
        n = self._check(self._l.snb_profile_pass(self._h, batch, arr, 512))
        return [(arr[i].name.decode(), arr[i].ms, arr[i].flops, arr[i].bytes) for i in range(n)]


class Pool:
    """One model replica per GPU behind one handle (snb_pool_*): replica 0 loads the weights, one NCCL broadcast installs
    them on the other GPUs, calls go to the least busy replica, batches are sharded contiguously."""

    def __init__(self, height: int, width: int, K: int, D: int, *, devices=None, max_batch: int = 1, task_num: int = 4,
                 precision: int = PREC_TC_F16X2, flags: int = 0, model_file: Optional[str] = None, weights: Optional[bytes] = None):
        self._l = lib()
        self._h = C.c_void_p()
        self._cbs = {}
        cfg = SnbConfig(C.sizeof(SnbConfig), height, width, K, D, max_batch, 0, task_num, precision, flags,
                        model_file.encode() if model_file else None, None, 0)
        self._wbuf = None
        if weights is not None:
            self._wbuf = C.create_string_buffer(weights, len(weights))
            cfg.weights = C.cast(self._wbuf, C.c_void_p)
            cfg.weights_bytes = len(weights)
        devs = (C.c_int32 * len(devices))(*devices) if devices else None
        r = self._l.snb_pool_create(C.byref(self._h), C.byref(cfg), devs, len(devices) if devices else 0)
        if r != SNB_OK:
            raise SnbError(r, (self._l.snb_pool_last_error(None) or b"").decode())
        self.H, self.W, self.K, self.D = height, width, K, D

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._l.snb_pool_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _check(self, r: int):
        if r < 0:
            raise SnbError(r, (self._l.snb_pool_last_error(self._h) or b"").decode())
        return r

    def size(self) -> int:
        return int(self._l.snb_pool_size(self._h))

    def infer(self, s8, out=None):
        B = s8.shape[0]
        if out is None:
            out = np.empty((B, 1, self.H, self.W), np.int32)
        self._check(self._l.snb_pool_infer(self._h, _ptr(s8), _ptr(out), B))
        return out

    def infer_async(self, x, out, nv12: bool = False, timeout_ms: int = -1):
        key = id(out)

        def _cb(user, status, stat):
            self._cbs.pop(key, None)

        cb = DONE_FN(_cb)
        self._cbs[key] = (cb, x, out)
        fn = self._l.snb_pool_infer_nv12_async if nv12 else self._l.snb_pool_infer_async
        r = fn(self._h, _ptr(x), _ptr(out), x.shape[0], cb, None, timeout_ms)
        if r < 0:
            self._cbs.pop(key, None)
            self._check(r)

    def wait_all(self):
        self._check(self._l.snb_pool_wait_all(self._h))

    def stat(self) -> dict:
        s = SnbPoolStat()
        self._check(self._l.snb_pool_get_stat(self._h, C.byref(s)))
        n = s.n_devices
        return {"n_devices": n, "devices": list(s.device[:n]), "calls": list(s.calls[:n]), "weight_bytes": int(s.weight_bytes),
                "broadcast_ms": float(s.broadcast_ms)}


def shard_range(n_pairs: int, world: int, rank: int):
    a, b = C.c_int64(), C.c_int64()
    r = lib().snb_shard_range(n_pairs, world, rank, C.byref(a), C.byref(b))
    if r < 0:
        raise SnbError(r, "snb_shard_range")
    return a.value, b.value


def weights_validate(blob: bytes, K: int = 0) -> None:
    """Raises SnbError(SNB_ERR_MODEL, reason) unless `blob` is a weight blob snb_create would accept."""
    buf = C.create_string_buffer(blob, len(blob))
    r = lib().snb_weights_validate(C.cast(buf, C.c_void_p), len(blob), K)
    if r < 0:
        raise SnbError(r, (lib().snb_last_error(None) or b"").decode())


def nv12_to_bgr(nv12: np.ndarray, w: int, h: int) -> np.ndarray:
    out = np.empty((h, w, 3), np.uint8)
    r = lib().snb_pre_nv12_to_bgr(_ptr(np.ascontiguousarray(nv12, np.uint8)), w, h, _ptr(out))
    if r < 0:
        raise SnbError(r, "snb_pre_nv12_to_bgr")
    return out


def jpeg_encode_nv12(nv12: np.ndarray, w: int, h: int, quality: int = 0) -> bytes:
    src = np.ascontiguousarray(nv12, np.uint8)
    dst = np.empty(w * h * 3 + 4096, np.uint8)
    n = lib().snb_jpeg_encode_nv12(_ptr(src), w, h, quality, _ptr(dst), dst.size)
    if n < 0:
        raise SnbError(int(n), "snb_jpeg_encode_nv12")
    if n > dst.size:                       # the call reports the size it needs and never writes past the buffer
        dst = np.empty(n, np.uint8)
        n = lib().snb_jpeg_encode_nv12(_ptr(src), w, h, quality, _ptr(dst), n)
    return dst[:n].tobytes()


def synthesize_weights(K: int, seed: int = 1234) -> bytes:
    """Seeded synthetic weight blob (SNB2WGT1) made by the library itself."""
    n = lib().snb_weights_synthesize(K, seed, None, 0)
    if n < 0:
        raise SnbError(int(n), "snb_weights_synthesize")
    buf = C.create_string_buffer(n)
    n2 = lib().snb_weights_synthesize(K, seed, C.cast(buf, C.c_void_p), n)
    if n2 != n:
        raise SnbError(int(n2), "snb_weights_synthesize")
    return buf.raw


# ---- host-side byte formats (no GPU involved) ---------------------------------------------------
def pre_split_nv12(frame: np.ndarray, h: int, w2: int):
    w = w2 // 2
    left = np.empty(h * 3 // 2 * w, np.uint8)
    right = np.empty_like(left)
    r = lib().snb_pre_split_nv12(_ptr(np.ascontiguousarray(frame, np.uint8)), h, w2, _ptr(left), _ptr(right))
    if r < 0:
        raise SnbError(r, "snb_pre_split_nv12")
    return left, right


def pre_yuv420_to_yuv444(buf: np.ndarray, w: int, h: int, correct_chroma: bool = False) -> np.ndarray:
    out = np.empty((3, h, w), np.uint8)
    r = lib().snb_pre_yuv420_to_yuv444(_ptr(np.ascontiguousarray(buf, np.uint8)), _ptr(out), w, h, int(correct_chroma))
    if r < 0:
        raise SnbError(r, "snb_pre_yuv420_to_yuv444")
    return out


def pre_cvt_nv12_to_tensor(left: np.ndarray, right: np.ndarray, w: int, h: int, correct_chroma: bool = False):
    out = np.empty((1, 6, h, w), np.int8)
    r = lib().snb_pre_cvt_nv12_to_tensor(_ptr(np.ascontiguousarray(left, np.uint8)),
                                         _ptr(np.ascontiguousarray(right, np.uint8)), w, h, int(correct_chroma), _ptr(out))
    if r < 0:
        raise SnbError(r, "snb_pre_cvt_nv12_to_tensor")
    return out


def pre_quantize(v: float, scale=0.0078125, zero_point=0.5, lo=-128.0, hi=127.0) -> int:
    return int(lib().snb_pre_quantize(v, scale, zero_point, lo, hi))


def post_pack(infer: np.ndarray, jpeg: bytes) -> bytes:
    infer = np.ascontiguousarray(infer, np.int32)
    dst = np.empty(infer.nbytes + len(jpeg), np.uint8)
    jb = np.frombuffer(jpeg, np.uint8) if len(jpeg) else None
    n = lib().snb_post_pack(_ptr(infer), infer.nbytes, _ptr(jb) if jb is not None else None, len(jpeg), _ptr(dst), dst.nbytes)
    if n < 0:
        raise SnbError(int(n), "snb_post_pack")
    return dst[:n].tobytes()


def post_parse_depth(q: np.ndarray, scale: float = 2.60443857769133e-06) -> np.ndarray:
    q = np.ascontiguousarray(q, np.int32)
    out = np.empty(q.shape, np.float32)
    r = lib().snb_post_parse_depth(_ptr(q), q.size, scale, _ptr(out))
    if r < 0:
        raise SnbError(r, "snb_post_parse_depth")
    return out
