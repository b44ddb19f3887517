#!/usr/bin/env python
"""Launch the small streaming kernels of the path at QuickNet-scale sizes (for `ncu -k regex:...`):
bmaxpool_kernel, unpack_kernel (LceDequantize), pack_generic_kernel (LceQuantize of int8 input /
ragged channel counts), pack_f32_flat_kernel. Development / evidence tool."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compute_engine_b200 import capi  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, 56, 56, 64, device="cuda", generator=g)
for _ in range(2):
    q = capi.quantize(x)                                   # pack_f32_flat_kernel
    xi = torch.randint(-128, 128, (B, 56, 56, 72), device="cuda", generator=g, dtype=torch.int16).to(torch.int8)
    qi = capi.quantize(xi, zero_point=3)                   # pack_generic_kernel<int8>, ragged channels
    xr = torch.randn(B, 56, 56, 40, device="cuda", generator=g)
    qr = capi.quantize(xr)                                 # pack_generic_kernel<float>, 40 channels
    p = capi.bmaxpool(q, (3, 3), (2, 2), capi.PADDING_SAME)   # bmaxpool_kernel
    d = capi.dequantize(q, 64)                             # unpack_kernel<float>
torch.cuda.synchronize()
print("done", q.shape, qi.shape, qr.shape, p.shape, d.shape)
