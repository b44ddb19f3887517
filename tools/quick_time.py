"""Quick device timing of the hot kernels (development aid; bench.py is the
contract). CUDA events on the current stream, L2 flushed between iterations."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compute_engine_b200 import capi  # noqa: E402


def time_fn(fn, iters=10, warmup=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    g = torch.Generator(device="cuda").manual_seed(0)
    for (hw, c) in [(56, 64), (28, 128), (14, 256), (7, 512)]:
        cw = c // 32
        filt = torch.randint(-2**31, 2**31 - 1, (c, 3, 3, cw), device="cuda", generator=g,
                             dtype=torch.int64).to(torch.int32)
        mul = torch.rand(c, device="cuda") + 0.1
        bias = torch.rand(c, device="cuda")
        d = capi.BconvDesc(B, hw, hw, c, 3, 3, c, 1, 1, 1, 1, 1, capi.PADDING_SAME, 1,
                           capi.ACT_RELU, capi.OUT_FLOAT, 1.0, 0)
        plan = capi.BConv2d(d, filt, mul, bias)
        xf = torch.randn((B, hw, hw, c), device="cuda", generator=g)
        xp = capi.quantize(xf)
        out = torch.empty((B, hw, hw, c), device="cuda")
        macs = B * hw * hw * c * 9 * c
        for name, fn, in_bytes in (("packed", lambda: plan(xp, out), xp.numel() * 4),
                                   ("f32in", lambda: plan(xf, out), xf.numel() * 4)):
            med, best = time_fn(fn)
            bytes_alg = in_bytes + filt.numel() * 4 + out.numel() * 4 + 8 * c
            print(json.dumps({"kernel": f"bconv {hw}x{hw}x{c} {name}", "batch": B,
                              "ms": round(med, 4), "ms_best": round(best, 4),
                              "binary_TOPS": round(2 * macs / med / 1e9, 1),
                              "word_ops_per_s_T": round(macs / 32 / med / 1e9, 3),
                              "alg_GBps": round(bytes_alg / med / 1e6, 1),
                              "img_per_s": round(B / med * 1e3, 1)}))
        import ctypes as C
        res = torch.randn((B, hw, hw, c), device="cuda", generator=g)
        pk = torch.empty((B, hw, hw, cw), device="cuda", dtype=torch.int32)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def fused():
            rc = capi.lib().lce_b200_bconv2d_run_fused(
                plan._h, C.c_void_p(xp.data_ptr()), C.c_void_p(res.data_ptr()), 0,
                C.c_void_p(out.data_ptr()), C.c_void_p(pk.data_ptr()), st)
            assert rc == 0
        med, best = time_fn(fused)
        print(json.dumps({"kernel": f"bconv {hw}x{hw}x{c} fused(+ADD+LceQuantize)", "batch": B,
                          "ms": round(med, 4), "ms_best": round(best, 4),
                          "word_ops_per_s_T": round(macs / 32 / med / 1e9, 3)}))
        med, best = time_fn(lambda: capi.quantize(xf, out=xp))
        print(json.dumps({"kernel": f"bsign_pack {hw}x{hw}x{c}", "ms": round(med, 4),
                          "alg_GBps": round((xf.numel() * 4 + xp.numel() * 4) / med / 1e6, 1)}))
        plan.close()
    for (M, N, Kb) in [(4096, 4096, 8192), (4096, 4096, 256), (1024, 1024, 2048),
                       (256, 256, 256), (4096, 256, 2304)]:
        Kw = Kb // 32
        A = torch.randint(-2**31, 2**31 - 1, (M, Kw), device="cuda", generator=g,
                          dtype=torch.int64).to(torch.int32)
        W = torch.randint(-2**31, 2**31 - 1, (N, Kw), device="cuda", generator=g,
                          dtype=torch.int64).to(torch.int32)
        gemm = capi.BGemm(W)
        out = torch.empty((M, N), dtype=torch.int32, device="cuda")
        med, best = time_fn(lambda: gemm(A, out))
        bytes_alg = (M + N) * Kw * 4 + M * N * 4
        print(json.dumps({"kernel": f"bgemm M{M} N{N} Kbits{Kb} int32-out", "ms": round(med, 4),
                          "ms_best": round(best, 4),
                          "binary_TOPS": round(2 * M * N * Kb / med / 1e9, 1),
                          "word_ops_per_s_T": round(M * N * Kw / med / 1e9, 3),
                          "alg_GBps": round(bytes_alg / med / 1e6, 1)}))
        gemm.close()


if __name__ == "__main__":
    main()
