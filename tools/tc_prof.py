"""Per-role cycle breakdown of the tcgen05 kernel (LCE_B200_TC_PROF=1) on the QuickNet stage
shapes and two BGEMM corners. Development tool."""
import os
import sys

os.environ["LCE_B200_TC_PROF"] = "1"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import ctypes as C  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

import lce_testlib as L  # noqa: E402
from compute_engine_b200 import capi  # noqa: E402

lib = capi.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for hw, c in ((56, 64), (28, 128), (14, 256), (7, 512)):
    case = L.make_bconv_case(1, 1, hw, hw, c, 3, 3, c, activation=L.ACT_RELU)
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.integers(-2**31, 2**31, (B, hw, hw, c // 32), dtype=np.int64).astype(np.int32)).cuda()
    res = torch.randn((B, hw, hw, c), device="cuda")
    d = capi.BconvDesc(*[getattr(case.desc, n) for n, _ in case.desc._fields_])
    d.batch = B
    plan = capi.BConv2d(d, case.filt, case.mul, case.bias, None)
    out = torch.empty((B, hw, hw, c), device="cuda")
    pk = torch.empty((B, hw, hw, c // 32), dtype=torch.int32, device="cuda")
    print(f"--- conv {hw}x{hw}x{c} plain", file=sys.stderr, flush=True)
    plan(x, out)
    torch.cuda.synchronize()
    print(f"--- conv {hw}x{hw}x{c} fused", file=sys.stderr, flush=True)
    capi._check(lib.lce_b200_bconv2d_run_fused(plan._h, C.c_void_p(x.data_ptr()), C.c_void_p(res.data_ptr()), 0,
                                              C.c_void_p(out.data_ptr()), C.c_void_p(pk.data_ptr()), None))
    torch.cuda.synchronize()
    plan.close()
for (M, N, K) in ((4096, 4096, 256), (4096, 4096, 8192)):
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randint(-2**31, 2**31 - 1, (M, K // 32), device="cuda", generator=g, dtype=torch.int64).to(torch.int32)
    W = torch.randint(-2**31, 2**31 - 1, (N, K // 32), device="cuda", generator=g, dtype=torch.int64).to(torch.int32)
    gemm = capi.BGemm(W)
    print(f"--- bgemm {M}x{N}x{K}", file=sys.stderr, flush=True)
    gemm(A)
    torch.cuda.synchronize()
    gemm.close()
