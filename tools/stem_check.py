#!/usr/bin/env python
"""Time the fused stem kernel (lce_b200_f32_stem_conv_dw) against the three kernels it replaces,
at QuickNet's benched shape. Development tool. Usage: stem_check.py [batch] [reps]"""
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


def p(t):
    return C.c_void_p(t.data_ptr())


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    lib = capi.lib()
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randint(-128, 128, (B, 224, 224, 3), generator=g, dtype=torch.int16).to(torch.int8).cuda()
    w1 = (torch.randn(16, 3, 3, 3, generator=g) * 0.4).cuda()
    b1 = torch.randn(16, generator=g).cuda()
    w2 = (torch.randn(1, 3, 3, 16, generator=g) * 0.4).cuda()
    b2 = torch.randn(16, generator=g).cuda()
    d1 = ConvDesc(B, 224, 224, 3, 3, 3, 16, 2, 2, 1, 1, 0, 1)
    d2 = ConvDesc(B, 112, 112, 16, 3, 3, 16, 2, 2, 1, 1, 0, 0)
    xf = torch.empty(B, 224, 224, 3, device="cuda")
    y1 = torch.empty(B, 112, 112, 16, device="cuda")
    want = torch.empty(B, 56, 56, 16, device="cuda")
    got = torch.empty(B, 56, 56, 16, device="cuda")
    scale, zp = C.c_double(4.0 / 127), C.c_int32(0)
    hw = [t.cpu().contiguous() for t in (w1, b1, w2, b2)]   # the fused kernel takes the filters by value

    def separate():
        assert lib.lce_b200_dequantize_affine(1, p(x), p(xf), C.c_int64(x.numel()), scale, zp, None) == 0
        assert lib.lce_b200_f32_conv2d(C.byref(d1), p(xf), p(w1), p(b1), p(y1), None) == 0
        assert lib.lce_b200_f32_depthwise_conv2d(C.byref(d2), p(y1), p(w2), p(b2), p(want), None) == 0

    def fused():
        assert lib.lce_b200_f32_stem_conv_dw(C.byref(d1), C.byref(d2), 1, p(x), scale, zp, p(hw[0]), p(hw[1]), p(hw[2]),
                                             p(hw[3]), p(got), None) == 0

    for name, fn in (("separate", separate), ("fused", fused)):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:9s} {e0.elapsed_time(e1) * 1000 / reps:8.1f} us", flush=True)
    print("bit-identical:", torch.equal(got.view(torch.int32), want.view(torch.int32)))


if __name__ == "__main__":
    main()
