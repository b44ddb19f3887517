"""ctypes binding of the C-ABI CUDA layer (include/lce_b200.h).

PyTorch is used only as plumbing: device memory (torch tensors), streams and
``torch.distributed``. Every function here goes through ``liblce_b200.so``; there
is no eager / CPU fallback -- a missing library or device raises ``LceError``.

The Python names mirror the reference's op surface (LCE/tflite/kernels):
``quantize`` = LceQuantize, ``dequantize`` = LceDequantize, ``BConv2d`` =
LceBconv2d (Init/Prepare/Eval as constructor / set_input_shape / __call__),
``bmaxpool`` = LceBMaxPool2d, ``BGemm`` = core/bgemm.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import build as _build

PADDING_SAME, PADDING_VALID = 0, 1
ACT_NONE, ACT_RELU, ACT_RELU_N1_TO_1, ACT_RELU6 = 0, 1, 2, 3
OUT_FLOAT, OUT_INT8, OUT_BITPACKED, OUT_RAW_ACC = 0, 1, 2, 3
T_FLOAT, T_INT8, T_BOOL = 0, 1, 2


class LceError(RuntimeError):
    pass


class BconvDesc(C.Structure):
    """``lce_bconv2d_desc`` (include/lce_b200_types.h)."""

    _fields_ = [(n, C.c_int32) for n in (
        "batch", "in_h", "in_w", "channels_in", "filter_h", "filter_w",
        "channels_out", "groups", "stride_h", "stride_w", "dilation_h",
        "dilation_w", "padding", "pad_value", "activation", "out_type")] + [
        ("out_scale", C.c_float), ("out_zero_point", C.c_int32)]


class BMaxPoolDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "batch", "in_h", "in_w", "channels_packed", "filter_h", "filter_w",
        "stride_h", "stride_w", "padding")]


class BgemmEpilogue(C.Structure):
    _fields_ = [("out_type", C.c_int32), ("clamp_min", C.c_int32),
                ("clamp_max", C.c_int32), ("multiplier", C.c_void_p),
                ("bias", C.c_void_p), ("thresholds", C.c_void_p)]


EXPORTS = [
    "lce_b200_abi_version", "lce_b200_last_error", "lce_b200_device_count",
    "lce_b200_quantize", "lce_b200_dequantize", "lce_b200_bmaxpool_out_shape",
    "lce_b200_bmaxpool", "lce_b200_bconv2d_out_shape", "lce_b200_bconv2d_create",
    "lce_b200_bconv2d_set_input_shape", "lce_b200_bconv2d_get_desc",
    "lce_b200_bconv2d_set_zero_padding_mode", "lce_b200_path_counts", "lce_b200_tc_debug",
    "lce_b200_bconv2d_run", "lce_b200_bconv2d_run_fused", "lce_b200_bconv2d_run_f32", "lce_b200_bconv2d_run_host",
    "lce_b200_bconv2d_destroy", "lce_b200_bgemm_create", "lce_b200_bgemm_run",
    "lce_b200_bgemm_destroy", "lce_b200_launch_count",
]

_lib = None


def lib():
    """Load liblce_b200.so (built in-tree). Fails loudly if it is missing."""
    global _lib
    if _lib is None:
        path = os.environ.get("LCE_B200_LIB") or _build.cuda_lib_path()   # override: A/B experiments
        if not os.path.exists(path):
            raise LceError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (there is no CPU fallback)")
        _lib = C.CDLL(path)
        _lib.lce_b200_last_error.restype = C.c_char_p
        _lib.lce_b200_launch_count.restype = C.c_uint64
        if _lib.lce_b200_abi_version() != 1:
            raise LceError("liblce_b200.so ABI version mismatch")
    return _lib


def _check(rc):
    if rc != 0:
        raise LceError(lib().lce_b200_last_error().decode() or f"error {rc}")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _dev(t, dtype=None):
    if not t.is_cuda:
        raise LceError("expected a CUDA tensor (no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise LceError(f"expected dtype {dtype}, got {t.dtype}")
    return t.contiguous()


def path_counts():
    """Launches per binary inner product so far: [tcgen05, legacy mma.sync int8, XOR + POPC]."""
    a = (C.c_uint64 * 3)()
    lib().lce_b200_path_counts(a)
    return [int(v) for v in a]


def launch_count():
    return int(lib().lce_b200_launch_count())


def cdiv(a, b):
    return (a + b - 1) // b


# ------------------------------------------------------------------------- #
def quantize(x: torch.Tensor, zero_point: int = 0, out: torch.Tensor | None = None):
    """LceQuantize: bitpack the last axis (float32 / int8 / bool) into int32 words."""
    x = _dev(x)
    if x.dtype == torch.float32:
        t = T_FLOAT
    elif x.dtype == torch.int8:
        t = T_INT8
    elif x.dtype == torch.bool:
        t = T_BOOL
    else:
        raise LceError(f"LceQuantize: unsupported input type {x.dtype}")
    cols = x.shape[-1]
    rows = x.numel() // cols if cols else 0
    if out is None:
        out = torch.empty(x.shape[:-1] + (cdiv(cols, 32),), dtype=torch.int32, device=x.device)
    _check(lib().lce_b200_quantize(C.c_int(t), _p(x), C.c_int64(rows), C.c_int64(cols),
                                   C.c_int32(zero_point), _p(out), _stream()))
    return out


def dequantize(packed: torch.Tensor, channels: int, dtype=torch.float32, scale=1.0,
               zero_point=0):
    """LceDequantize: unpack int32 words into +-1 float / int8 / bool."""
    packed = _dev(packed, torch.int32)
    t = {torch.float32: T_FLOAT, torch.int8: T_INT8, torch.bool: T_BOOL}.get(dtype)
    if t is None:
        raise LceError(f"LceDequantize: unsupported output type {dtype}")
    if packed.shape[-1] != cdiv(channels, 32):
        raise LceError("LceDequantize: packed channels do not match")
    rows = packed.numel() // packed.shape[-1] if packed.shape[-1] else 0
    out = torch.empty(packed.shape[:-1] + (channels,), dtype=dtype, device=packed.device)
    _check(lib().lce_b200_dequantize(C.c_int(t), _p(packed), C.c_int64(rows),
                                     C.c_int64(channels), C.c_float(scale),
                                     C.c_int32(zero_point), _p(out), _stream()))
    return out


def bmaxpool(x: torch.Tensor, filter_hw, stride_hw, padding=PADDING_SAME):
    """LceBMaxPool2d on a bitpacked NHWC tensor."""
    x = _dev(x, torch.int32)
    b, h, w, c = x.shape
    d = BMaxPoolDesc(b, h, w, c, filter_hw[0], filter_hw[1], stride_hw[0], stride_hw[1], padding)
    oh, ow = C.c_int(), C.c_int()
    _check(lib().lce_b200_bmaxpool_out_shape(C.byref(d), C.byref(oh), C.byref(ow)))
    out = torch.empty((b, oh.value, ow.value, c), dtype=torch.int32, device=x.device)
    _check(lib().lce_b200_bmaxpool(C.byref(d), _p(x), _p(out), _stream()))
    return out


class BConv2d:
    """LceBconv2d plan: constructor = Init + Prepare + OneTimeSetup
    (LCE/tflite/kernels/bconv2d.cc:85-392), ``__call__`` = Eval (:551-564)."""

    def __init__(self, desc: BconvDesc, filt, post_mul=None, post_bias=None, thresholds=None):
        self._h = C.c_void_p()
        self.desc = desc
        keep = []

        def ptr(t, dtype):
            if t is None:
                return C.c_void_p(0)
            if isinstance(t, torch.Tensor):
                t = t.to(dtype).contiguous()
                keep.append(t)
                return C.c_void_p(t.data_ptr())
            import numpy as np
            a = np.ascontiguousarray(t, {torch.int32: np.int32, torch.float32: np.float32}[dtype])
            keep.append(a)
            return a.ctypes.data_as(C.c_void_p)

        rc = lib().lce_b200_bconv2d_create(
            C.byref(desc), ptr(filt, torch.int32), ptr(post_mul, torch.float32),
            ptr(post_bias, torch.float32), ptr(thresholds, torch.int32), C.byref(self._h))
        _check(rc)

    def set_input_shape(self, batch, in_h, in_w):
        _check(lib().lce_b200_bconv2d_set_input_shape(self._h, batch, in_h, in_w))
        self.desc.batch, self.desc.in_h, self.desc.in_w = batch, in_h, in_w

    def set_zero_padding_mode(self, mode):
        """0: the reference kernel's integers (reference.h:100-103); 1: the optimised kernels'
        float correction (zero_padding_correction.h) -- see include/lce_b200_types.h."""
        _check(lib().lce_b200_bconv2d_set_zero_padding_mode(self._h, int(mode)))

    def out_shape(self):
        d = BconvDesc()
        oh, ow = C.c_int(), C.c_int()
        _check(lib().lce_b200_bconv2d_get_desc(self._h, C.byref(d), C.byref(oh), C.byref(ow)))
        last = cdiv(d.channels_out, 32) if d.out_type == OUT_BITPACKED else d.channels_out
        return (d.batch, oh.value, ow.value, last)

    def out_dtype(self):
        return {OUT_FLOAT: torch.float32, OUT_INT8: torch.int8,
                OUT_BITPACKED: torch.int32}[self.desc.out_type]

    def _alloc(self, like, out):
        if out is None:
            out = torch.empty(self.out_shape(), dtype=self.out_dtype(), device=like.device)
        return out

    def __call__(self, x: torch.Tensor, out: torch.Tensor | None = None):
        """x: bitpacked NHWC int32, or float32 NHWC (LceQuantize fused in front)."""
        x = _dev(x)
        if tuple(x.shape[:3]) != (self.desc.batch, self.desc.in_h, self.desc.in_w):
            self.set_input_shape(*x.shape[:3])
        out = self._alloc(x, out)
        if x.dtype == torch.float32:
            if x.shape[3] != self.desc.channels_in:
                raise LceError("LceBconv2d: float input channels do not match channels_in")
            _check(lib().lce_b200_bconv2d_run_f32(self._h, _p(x), _p(out), _stream()))
        elif x.dtype == torch.int32:
            if x.shape[3] != cdiv(self.desc.channels_in, 32):
                raise LceError("LceBconv2d: packed input channels do not match channels_in")
            _check(lib().lce_b200_bconv2d_run(self._h, _p(x), _p(out), _stream()))
        else:
            raise LceError(f"LceBconv2d: unsupported input type {x.dtype}")
        return out

    def run_host(self, x_host, out_host):
        """Host-buffer call (numpy arrays): H2D + kernel + D2H, synchronised."""
        _check(lib().lce_b200_bconv2d_run_host(
            self._h, x_host.ctypes.data_as(C.c_void_p), out_host.ctypes.data_as(C.c_void_p)))
        return out_host

    def close(self):
        if self._h:
            lib().lce_b200_bconv2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BGemm:
    """Binary GEMM plan: out[M,N] = epilogue(sum_k popc(A[m,k] ^ W[n,k]))."""

    def __init__(self, W, out_type=OUT_RAW_ACC, clamp=(0, 2**31 - 1), multiplier=None,
                 bias=None, thresholds=None):
        import numpy as np
        self._h = C.c_void_p()
        self.out_type = out_type
        if isinstance(W, torch.Tensor):
            W = W.to(torch.int32).contiguous()
            wp = C.c_void_p(W.data_ptr())
        else:
            W = np.ascontiguousarray(W, np.int32)
            wp = W.ctypes.data_as(C.c_void_p)
        self.N, self.Kw = W.shape
        keep = [W]

        def hp(a, dt):
            if a is None:
                return None
            if isinstance(a, torch.Tensor):
                a = a.detach().cpu().numpy()
            a = np.ascontiguousarray(a, dt)
            keep.append(a)
            return a.ctypes.data_as(C.c_void_p)

        ep = BgemmEpilogue(out_type, clamp[0], clamp[1], hp(multiplier, np.float32),
                           hp(bias, np.float32), hp(thresholds, np.int32))
        _check(lib().lce_b200_bgemm_create(C.c_int(self.N), C.c_int(self.Kw), wp, C.byref(ep),
                                           C.byref(self._h)))

    def __call__(self, A: torch.Tensor, out: torch.Tensor | None = None):
        A = _dev(A, torch.int32)
        M = A.shape[0]
        if A.shape[1] != self.Kw:
            raise LceError("BGemm: K mismatch")
        if out is None:
            if self.out_type == OUT_BITPACKED:
                out = torch.empty((M, cdiv(self.N, 32)), dtype=torch.int32, device=A.device)
            else:
                dt = {OUT_FLOAT: torch.float32, OUT_INT8: torch.int8,
                      OUT_RAW_ACC: torch.int32}[self.out_type]
                out = torch.empty((M, self.N), dtype=dt, device=A.device)
        _check(lib().lce_b200_bgemm_run(self._h, C.c_int64(M), _p(A), _p(out), _stream()))
        return out

    def close(self):
        if self._h:
            lib().lce_b200_bgemm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
