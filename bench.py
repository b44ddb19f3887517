#!/usr/bin/env python
"""bench.py -- the measurement contract.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input. Prints ONE
JSON line on rank 0. See DESIGN.md section "Measurement" for every field.

Workload (config.workload):
  quicknet_graph_b256       the full QuickNet .tflite graph through the graph host
                            (used when compute_engine_b200.graph is available)
  quicknet_bconv_stack_b256 the 16 LceQuantize->LceBconv2d layers of QuickNet
                            (4 per stage, 3x3 s1 SAME one-padding, fused ReLU, float in ->
                            float out), batch 256 per GPU -- the path BASELINE.json names.
The `--impl reference` arm times the reference's own CPU kernels (oracle/_ref: its
headers compiled by oracle/Makefile; else the C port) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

STAGES = [(56, 64), (28, 128), (14, 256), (7, 512)]   # (H=W, C) of QuickNet's 4 sections
LAYERS_PER_STAGE = 4


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU")
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(p.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- #
# synthetic workload definition (shared by both arms; seeded)
# --------------------------------------------------------------------------- #
def make_stack_weights(seed=0):
    """Random +-1 filters (bitpacked OHWI) and BatchNorm-like multiplier / bias for
    the 16 binary convolutions of QuickNet."""
    rng = np.random.default_rng(seed)
    layers = []
    for (hw, c) in STAGES:
        for _ in range(LAYERS_PER_STAGE):
            filt = rng.integers(-2**31, 2**31, (c, 3, 3, c // 32), dtype=np.int64).astype(np.int32)
            mul = rng.uniform(0.01, 1.5, c).astype(np.float32) / np.float32(9 * c)
            bias = rng.uniform(-1.0, 1.0, c).astype(np.float32)
            layers.append({"hw": hw, "c": c, "filter": filt, "mul": mul, "bias": bias})
    return layers


def make_stage_inputs(batch, seed):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((batch, hw, hw, c), dtype=np.float32) for (hw, c) in STAGES]


def layer_alg_bytes(batch, hw, c, s_in=4):
    """SURVEY 8(d) algorithmic bytes of one fused (quantize + bconv) layer."""
    return batch * hw * hw * c * s_in + c * 9 * c // 8 + batch * hw * hw * c * 4 + 8 * c


def layer_word_ops(batch, hw, c):
    return batch * hw * hw * c * 9 * (c // 32)


# --------------------------------------------------------------------------- #
# reference arm: the reference's own CPU kernels on the host cores
# --------------------------------------------------------------------------- #
def run_reference_stack(batch, steps, warmup, sample_images=None):
    """Times LceQuantize + LceBconv2d for the 16 layers with the reference's own
    code: bitpack_matrix + indirect-BGEMM Kernel4x2Portable (the fastest path the
    reference has on x86) from oracle/_ref when present (kind 'reference'), else the
    C port in oracle/ (kind 'port'). Images are spread over all host cores."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import lce_testlib as L  # the cpu-baseline leg is the one place bench.py may use oracle/
    cores = os.cpu_count() or 1
    impl = "ref" if L.load_ref() is not None else "oracle"
    kind = "reference" if impl == "ref" else "port"
    n_img = sample_images or batch
    layers = make_stack_weights(0)
    inputs = [x[:n_img] for x in make_stage_inputs(max(n_img, 1), 1)]

    def one_step():
        li = 0
        for s, (hw, c) in enumerate(STAGES):
            x = inputs[s]
            for _ in range(LAYERS_PER_STAGE):
                lay = layers[li]
                li += 1
                packed = L.quantize(x, impl=impl)
                d = L.BconvDesc(n_img, hw, hw, c, 3, 3, c, 1, 1, 1, 1, 1, L.PADDING_SAME, 1,
                                L.ACT_RELU, L.OUT_FLOAT, 1.0, 0)
                x = L.bconv2d(d, packed, lay["filter"], lay["mul"], lay["bias"], impl=impl,
                              kind=1, threads=cores)
        return x

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"images_per_s": n_img / dt, "ms_per_step": dt * 1e3, "kind": kind, "cores": cores,
            "sample": f"{n_img} images x 16 layers (bitpack + "
                      f"{'Kernel4x2Portable indirect BGEMM' if impl == 'ref' else 'C port'}), "
                      f"{steps} timed steps after {warmup} warm-up, one image per task on "
                      f"{cores} threads"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_img = min(args.batch, 256)
    r = run_reference_stack(args.batch, max(1, min(args.steps, 5)), min(args.warmup, 1), n_img)
    line = {
        "impl": "reference", "metric": "quicknet_binary_conv_stack_images_per_sec",
        "value": r["images_per_s"], "unit": "images/s", "n_gpus": args.gpus,
        "steps": max(1, min(args.steps, 5)), "warmup": min(args.warmup, 1),
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 xor-popcount, f32 epilogue", "data": "synthetic",
        "config": {"workload": "quicknet_bconv_stack_b256", "batch_per_step": n_img,
                   "layers": 16, "note": "reference CPU kernels; rank 0 only"},
        "cpu_baseline": {"value": r["images_per_s"], "unit": "images/s", "cores": r["cores"],
                         "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["images_per_s"], "unit": "images/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- #
# B200 arm
# --------------------------------------------------------------------------- #
def main_b200(args):
    import torch
    import torch.distributed as dist
    from compute_engine_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    capi.lib()

    B = args.batch
    # ---- load: rank 0 owns the model; ONE broadcast of the packed weights -------
    from compute_engine_b200 import parallel
    layers = make_stack_weights(0) if rank == 0 else None
    if rank == 0:
        layers = [{"filter": l["filter"], "mul": l["mul"], "bias": l["bias"]} for l in layers]
    dev_layers, blob = parallel.broadcast_model(layers, dev)   # the only collective; none per step
    shapes = [(c, hw) for (hw, c) in STAGES for _ in range(LAYERS_PER_STAGE)]
    plans = []
    for (c, hw), lay in zip(shapes, dev_layers):
        d = capi.BconvDesc(B, hw, hw, c, 3, 3, c, 1, 1, 1, 1, 1, capi.PADDING_SAME, 1,
                           capi.ACT_RELU, capi.OUT_FLOAT, 1.0, 0)
        plans.append(capi.BConv2d(d, lay["filter"], lay["mul"], lay["bias"]))

    # ---- per-rank shard of the batch (images are independent) -------------------
    host_in = [torch.from_numpy(x).pin_memory() for x in make_stage_inputs(B, 100 + rank)]
    dev_in = [x.to(dev, non_blocking=True) for x in host_in]
    bufs = [[torch.empty((B, hw, hw, c), device=dev) for _ in range(2)] for (hw, c) in STAGES]
    packed = [torch.empty((B, hw, hw, c // 32), dtype=torch.int32, device=dev) for (hw, c) in STAGES]
    host_out = [torch.empty((B, hw, hw, c), dtype=torch.float32).pin_memory() for (hw, c) in STAGES]
    torch.cuda.synchronize()

    ev_pairs = []

    def step(inputs, record=False):
        li = 0
        outs = []
        for s in range(len(STAGES)):
            x = inputs[s]
            for k in range(LAYERS_PER_STAGE):
                capi.quantize(x, out=packed[s])
                y = bufs[s][k & 1]
                if record:
                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                    plans[li](packed[s], out=y)
                    e1.record()
                    ev_pairs.append((e0, e1))
                else:
                    plans[li](packed[s], out=y)
                x = y
                li += 1
            outs.append(x)
        return outs

    def step_e2e():
        for d, h in zip(dev_in, host_in):
            d.copy_(h, non_blocking=True)
        outs = step(dev_in)
        for h, o in zip(host_out, outs):
            h.copy_(o, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    W = max(args.warmup, 3)
    K = args.steps
    for _ in range(W):
        step(dev_in)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = capi.launch_count()
    total_ms = timed(lambda: step(dev_in, record=True), K)
    launches = capi.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # dominant kernel: the binary conv, timed live over the timed region
    torch.cuda.synchronize()
    conv_ms = sum(a.elapsed_time(b) for a, b in ev_pairs)
    n_conv = len(ev_pairs)
    ev_pairs.clear()

    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, K)

    if world > 1:
        t = torch.tensor([float(launches)], device=dev)
        dist.all_reduce(t)
        launches = int(t.item())

    if rank == 0:
        hbm_peak, peak_src, sm_max = measured_peaks()
        ms_per_step = total_ms / K
        alg_bytes = sum(layer_alg_bytes(B, hw, c, s_in=0.125) for (hw, c) in STAGES) * LAYERS_PER_STAGE
        word_ops = sum(layer_word_ops(B, hw, c) for (hw, c) in STAGES) * LAYERS_PER_STAGE
        conv_s = conv_ms * 1e-3 / K          # bconv kernel time per step (this rank)
        achieved = alg_bytes / conv_s / 1e9
        popc_peak = 148 * 16 * sm_max * 1e6  # measured 15.98 POPC/clk/SM (profiles/r01_microbench_pipes.jsonl)
        in_bytes = sum(x.numel() * 4 for x in host_in)
        out_bytes = sum(x.numel() * 4 for x in host_out)
        line = {
            "metric": "quicknet_binary_conv_stack_images_per_sec",
            "value": B * world / (ms_per_step * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 xor-popcount, f32 epilogue", "data": "synthetic",
            "config": {"workload": "quicknet_bconv_stack_b256", "batch_per_gpu": B,
                       "global_batch": B * world, "layers": 16,
                       "stages": [f"{hw}x{hw}x{c}" for hw, c in STAGES],
                       "parallelism": f"dp{world} (batch-sharded, weights broadcast once)",
                       "l2": "inputs+activations per step (>1 GB) exceed the 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": B * world / (e2e_ms / K * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                    "ms_per_step": e2e_ms / K},
            "gpu_launches": launches,
            "roofline": {"kernel": "lce::bconv_kernel<V,float>", "bound": "hbm",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "peak_source": peak_src, "traffic": None,
                         "launches_timed": n_conv, "avg_launch_ms": conv_ms / max(n_conv, 1),
                         "share_of_step": conv_ms / total_ms,
                         "int_pipe": {"achieved_word_ops_per_s": word_ops / conv_s,
                                      "peak_popc_per_s": popc_peak,
                                      "frac": word_ops / conv_s / popc_peak,
                                      "note": "XOR+POPC formulation is POPC-pipe bound: 16 POPC/clk/SM measured"}},
        }
        if not args.no_cpu_baseline and world == 1:
            r = run_reference_stack(B, 2, 1, min(B, 256))
            line["cpu_baseline"] = {"value": r["images_per_s"], "unit": "images/s",
                                    "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
