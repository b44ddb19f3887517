#!/usr/bin/env python
"""Time the fused MAX_POOL_2D 2x2/s1 + DEPTHWISE_CONV_2D 3x3/s2 kernel (QuickNet's blur-pool
pairs) at the benched shapes. Development tool. Usage: pool_check.py [batch] [reps]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compute_engine_b200 import capi  # noqa: E402


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("batch", "in_h", "in_w", "in_c", "filter_h", "filter_w", "out_c",
                                         "stride_h", "stride_w", "dilation_h", "dilation_w", "padding",
                                         "activation")]


class PoolDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("batch", "in_h", "in_w", "channels", "filter_h", "filter_w", "stride_h",
                                         "stride_w", "padding", "activation")]


def p(t):
    return C.c_void_p(t.data_ptr())


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    lib = capi.lib()
    g = torch.Generator(device="cuda").manual_seed(0)
    for hw, c in ((56, 64), (28, 128), (14, 256)):
        x = torch.randn(B, hw, hw, c, device="cuda", generator=g)
        w = torch.randn(1, 3, 3, c, device="cuda", generator=g)
        b = torch.randn(c, device="cuda", generator=g)
        pd = PoolDesc(B, hw, hw, c, 2, 2, 1, 1, 1, 0)                      # VALID
        dd = ConvDesc(B, hw - 1, hw - 1, c, 3, 3, c, 2, 2, 1, 1, 0, 0)    # SAME on the pooled map
        out = torch.empty(B, hw // 2, hw // 2, c, device="cuda")

        def call():
            assert lib.lce_b200_f32_maxpool2x2_depthwise3x3(C.byref(pd), C.byref(dd), p(x), p(w), p(b), p(out), None) == 0, \
                lib.lce_b200_last_error().decode()

        call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / reps
        gb = (x.numel() + out.numel()) * 4 / 1e9
        print(f"{hw}x{hw}x{c}: {us:7.1f} us  {gb / (us * 1e-6):6.0f} GB/s  checksum {out.double().sum().item():.6e}", flush=True)


if __name__ == "__main__":
    main()
