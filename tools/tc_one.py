"""Run one binary-convolution shape a few times (for `ncu -k regex:bconv_tc`). Development tool.
  python tools/tc_one.py HW C [fused|plain] [batch]"""
import ctypes as C
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import lce_testlib as L  # noqa: E402
from compute_engine_b200 import capi  # noqa: E402

hw, c = int(sys.argv[1]), int(sys.argv[2])
fused = len(sys.argv) > 3 and sys.argv[3] == "fused"
B = int(sys.argv[4]) if len(sys.argv) > 4 else 256
case = L.make_bconv_case(1, 1, hw, hw, c, 3, 3, c, activation=L.ACT_RELU)
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.integers(-2**31, 2**31, (B, hw, hw, c // 32), dtype=np.int64).astype(np.int32)).cuda()
res = torch.randn((B, hw, hw, c), device="cuda")
d = capi.BconvDesc(*[getattr(case.desc, n) for n, _ in case.desc._fields_])
d.batch = B
plan = capi.BConv2d(d, case.filt, case.mul, case.bias, None)
out = torch.empty((B, hw, hw, c), device="cuda")
pk = torch.empty((B, hw, hw, c // 32), dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_()
    if fused:
        capi._check(capi.lib().lce_b200_bconv2d_run_fused(plan._h, C.c_void_p(x.data_ptr()), C.c_void_p(res.data_ptr()), 0,
                                                          C.c_void_p(out.data_ptr()), C.c_void_p(pk.data_ptr()), None))
    else:
        plan(x, out)
torch.cuda.synchronize()
print("done", hw, c, fused, B)
