"""Time one BGEMM point (development aid): python tools/bgemm_point.py M N K_bits [noflush]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compute_engine_b200 import capi  # noqa: E402

M, N, Kb = (int(a) for a in sys.argv[1:4])
noflush = len(sys.argv) > 4 and sys.argv[4] == "noflush"
g = torch.Generator(device="cuda").manual_seed(0)
Kw = Kb // 32
A = torch.randint(-2**31, 2**31 - 1, (M, Kw), device="cuda", generator=g, dtype=torch.int64).to(torch.int32)
W = torch.randint(-2**31, 2**31 - 1, (N, Kw), device="cuda", generator=g, dtype=torch.int64).to(torch.int32)
gemm = capi.BGemm(W)
out = torch.empty((M, N), dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    gemm(A, out)
ts = []
for _ in range(15):
    if not noflush:
        flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gemm(A, out); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
# back-to-back launches: amortises the launch gap of a single event pair
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    gemm(A, out)
e1.record()
torch.cuda.synchronize()
ms = float(np.median(ts))
alg = (M + N) * Kw * 4 + M * N * 4
print(json.dumps({"M": M, "N": N, "K_bits": Kb, "flush": not noflush, "ms": round(ms, 5),
                  "ms_min": round(min(ts), 5), "ms_back_to_back": round(e0.elapsed_time(e1) / 20, 5),
                  "hbm_frac": round(alg / ms / 1e6 / 6571.2, 4)}))
