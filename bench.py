#!/usr/bin/env python
"""bench.py -- the measurement contract.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload W]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the workload over one batch of synthetic input; rank 0 prints
ONE JSON line. Every field is explained in DESIGN.md section 6. At N = 1 the graph workloads
are measured twice in the same run: with the default kernel selection (`value`, `roofline`)
and with every LceBconv2d on the XOR + POPC kernel (`xor_popc_path`, LCE_B200_BCONV_IMMA=0).

Workloads (config.workload):
  quicknet         (default; BASELINE.json configs[1]) the full QuickNet `.tflite` graph,
                   batch 256 per GPU, through the graph host + custom-op registrations
  quicknet_large   configs[2] (batch 1024 sharded over the GPUs: use --batch 128 --gpus 8)
  birealnet18      configs[3] (--batch 512)
  bconv_stack      the 16 LceQuantize->LceBconv2d layers of QuickNet alone
  bgemm_sweep      configs[4]: prints one extra JSON object per (M, N, K) to stderr
The `--impl reference` arm times the reference's own CPU kernels (oracle/_ref: its headers
compiled by oracle/Makefile; LCE ops: bitpack_matrix + Kernel4x2Portable indirect BGEMM)
with PyTorch-CPU fp32 standing in for TFLite's float builtins, on the same workload.
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
DEFAULT_BATCH = {"quicknet": 256, "quicknet_large": 128, "birealnet18": 512, "bconv_stack": 256,
                 "bgemm_sweep": 1}
METRIC = {"quicknet": "quicknet_images_per_sec", "quicknet_large": "quicknet_large_images_per_sec",
          "birealnet18": "birealnet18_images_per_sec",
          "bconv_stack": "quicknet_binary_conv_stack_images_per_sec",
          "bgemm_sweep": "bgemm_binary_tops"}


# Libraries (NCCL's version banner, ...) may write to fd 1; the contract is ONE JSON line on
# stdout, so fd 1 is pointed at stderr for the whole run and the result goes to the saved fd.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="quicknet", choices=sorted(DEFAULT_BATCH))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (0 = workload default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    if a.batch <= 0:
        a.batch = DEFAULT_BATCH[a.workload]
    return a


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
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            time.sleep(0.25)          # let the first samples arrive before the timed region
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.05):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- #
# synthetic model / data (seeded; identical for both arms)
# --------------------------------------------------------------------------- #
def build_model_bytes(workload):
    from compute_engine_b200 import zoo
    return {"quicknet": zoo.quicknet, "quicknet_large": zoo.quicknet_large,
            "birealnet18": zoo.birealnet18}[workload](batch=1, seed=0)


def make_images(batch, seed):
    return np.random.default_rng(seed).standard_normal((batch, 224, 224, 3), dtype=np.float32)


def make_stack_weights(seed=0):
    rng = np.random.default_rng(seed)
    layers = []
    for (hw, c) in STAGES:
        for _ in range(LAYERS_PER_STAGE):
            filt = rng.integers(-2**31, 2**31, (c, 3, 3, c // 32), dtype=np.int64).astype(np.int32)
            mul = rng.uniform(0.01, 1.5, c).astype(np.float32) / np.float32(9 * c)
            bias = rng.uniform(-1.0, 1.0, c).astype(np.float32)
            layers.append({"filter": filt, "mul": mul, "bias": bias})
    return layers


def make_stage_inputs(batch, seed):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((batch, hw, hw, c), dtype=np.float32) for (hw, c) in STAGES]


def measured_bf16_peak():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f).get("bf16_tflops"))
    except Exception:
        return None


def ncu_traffic_per_launch(fname="r01_ncu_bconv_v8_summary.csv"):
    """Mean dram__bytes_read + dram__bytes_write per lce::bconv_kernel launch from the committed
    `ncu --set full` capture of this same command (profiles/r01_ncu_bconv_v8_summary.csv:
    the 16 LceBconv2d launches of one QuickNet step), or None."""
    import csv
    path = os.path.join(REPO, "profiles", fname)
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
        tot = [float(r[ir]) * scale.get(units[ir], 1.0) + float(r[iw]) * scale.get(units[iw], 1.0)
               for r in rows[2:]]
        return sum(tot) / len(tot) if tot else None
    except Exception:
        return None


def bconv_alg_bytes(in_shape, filt_shape, out_shape, out_itemsize=4):
    """packed input + packed filter + output + multiplier/bias."""
    return (int(np.prod(in_shape)) * 4 + int(np.prod(filt_shape)) * 4 +
            int(np.prod(out_shape)) * out_itemsize + 8 * filt_shape[0])


def bconv_word_ops(out_shape, filt_shape):
    return int(np.prod(out_shape[:3])) * filt_shape[0] * int(np.prod(filt_shape[1:]))


# --------------------------------------------------------------------------- #
# reference arm: the reference's own CPU kernels on the host cores
# --------------------------------------------------------------------------- #
def run_reference(workload, n_img, steps, warmup):
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import lce_testlib as L          # the cpu-baseline leg is where bench.py may use oracle/
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    impl = "ref" if L.load_ref() is not None else "oracle"
    kind = "reference" if impl == "ref" else "port"
    if workload == "bconv_stack":
        layers = make_stack_weights(0)
        inputs = [x[:n_img] for x in make_stage_inputs(n_img, 1)]

        def one_step():
            li = 0
            for s, (hw, c) in enumerate(STAGES):
                x = inputs[s]
                for _ in range(LAYERS_PER_STAGE):
                    lay = layers[li]; li += 1
                    d = L.BconvDesc(n_img, hw, hw, c, 3, 3, c, 1, 1, 1, 1, 1, L.PADDING_SAME, 1,
                                    L.ACT_RELU, L.OUT_FLOAT, 1.0, 0)
                    x = L.bconv2d(d, L.quantize(x, impl=impl), lay["filter"], lay["mul"],
                                  lay["bias"], impl=impl, kind=1, threads=cores)
        what = "16 layers (bitpack + Kernel4x2Portable indirect BGEMM)"
    else:
        import tflite_ref as R
        model = R.parse(build_model_bytes(workload))
        x = make_images(n_img, 1)
        bk = 1 if impl == "ref" else 0

        def one_step():
            R.run(model, [x], threads=cores, lce_impl=impl, bconv_kind=bk)
        what = ("full graph: LCE ops = reference bitpack + Kernel4x2Portable indirect BGEMM, float "
                "builtins = PyTorch CPU fp32 (TFLite's own builtins cannot be built offline)")
    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"images_per_s": n_img / dt, "ms_per_step": dt * 1e3, "kind": kind, "cores": cores,
            "sample": f"{n_img} images per step, {steps} timed steps after {warmup} warm-up, "
                      f"{what}, one image per task on {cores} host threads"}


def cpu_baseline_subprocess(workload, batch):
    """The reference arm in its own process (its OpenMP settings must be in place before torch is
    imported; this process has long since imported it): bounded sample, 2 timed steps."""
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference",
                              "--workload", workload, "--batch", str(batch), "--steps", "2",
                              "--warmup", "1"], env=env, capture_output=True, text=True,
                             timeout=900, check=True).stdout.strip().splitlines()[-1]
        ref = json.loads(out)
        cb = dict(ref["cpu_baseline"])
        cb["value"], cb["unit"] = ref["value"], ref["unit"]
        return cb
    except Exception as e:  # never lose the GPU line over the baseline
        return {"value": None, "unit": "images/s", "error": str(e)[:200]}


def main_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    # The LCE kernels run on a pthread pool, the float builtins on torch's OpenMP pool. With the
    # default active wait policy torch's idle workers spin on every core and starve the LCE
    # threads (measured on the 128-core box: 15.7 -> 134.8 images/s just from this setting).
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
    os.environ.setdefault("KMP_BLOCKTIME", "0")
    if args.workload == "bgemm_sweep":
        emit({"impl": "reference", "unavailable": "bgemm_sweep has no reference arm"})
        return
    steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
    # one image per task: give every host thread at least one image
    n_img = min(args.batch, max(64, os.cpu_count() or 1) if args.workload != "bconv_stack" else 256)
    r = run_reference(args.workload, n_img, steps, warm)
    emit({
        "impl": "reference", "metric": METRIC[args.workload], "value": r["images_per_s"],
        "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 xor-popcount + f32", "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_step": n_img,
                   "note": "reference CPU kernels on the host cores; rank 0 only"},
        "cpu_baseline": {"value": r["images_per_s"], "unit": "images/s", "cores": r["cores"],
                         "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["images_per_s"], "unit": "images/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


# --------------------------------------------------------------------------- #
# B200 arm
# --------------------------------------------------------------------------- #
class Dist:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_ms(self, ms):
        t = self.torch.tensor([ms], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v):
        t = self.torch.tensor([float(v)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t)
        return float(t.item())

    def timed(self, fn, steps, sync, stream=None):
        """K steps bracketed by barrier + synchronize; device time from CUDA events on the
        launching stream; max over ranks."""
        torch = self.torch
        self.barrier()
        sync()
        stream = stream or torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall0 = time.time()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        sync()
        self.barrier()
        return self.max_ms(e0.elapsed_time(e1)), wall0, time.time()


def graph_workload(args, D):
    """Full `.tflite` graph through the graph host (custom-op registrations)."""
    torch = D.torch
    from compute_engine_b200 import capi, host as H
    B = args.batch
    # ---- load: rank 0 owns the model file; ONE broadcast of the flatbuffer ---------
    if D.rank == 0:
        blob = np.frombuffer(build_model_bytes(args.workload), np.uint8)
        n = torch.tensor([blob.size], device=D.dev, dtype=torch.int64)
    else:
        n = torch.zeros(1, device=D.dev, dtype=torch.int64)
    if D.world > 1:
        D.dist.broadcast(n, src=0)
    model = torch.empty(int(n.item()), dtype=torch.uint8, device=D.dev)
    if D.rank == 0:
        model.copy_(torch.from_numpy(blob.copy()))
    if D.world > 1:
        D.dist.broadcast(model, src=0)     # packed weights + graph; the only collective
    model_bytes = model.cpu().numpy().tobytes()
    g = H.HostGraph.from_tflite(model_bytes, device_arena=True)
    fused = 0 if os.environ.get("LCE_NO_FUSION") else g.fuse_all()
    t_in, t_out = g.inputs()[0], g.outputs()[0]
    g.resize_input(t_in, (B, 224, 224, 3))
    g.allocate_tensors()
    gs = torch.cuda.ExternalStream(g.stream())

    host_in = torch.from_numpy(make_images(B, 100 + D.rank)).pin_memory()
    host_out = torch.empty(g.shape(t_out), dtype=torch.float32).pin_memory()
    in_bytes, out_bytes = host_in.numel() * 4, host_out.numel() * 4
    g.write_ptr(t_in, host_in.data_ptr(), in_bytes)          # inputs resident in HBM
    g.synchronize()
    sync = g.synchronize

    # ---- (1) value: CUDA-graph replay, inputs resident -----------------------------
    g.invoke()                                               # eager: plans (weight tiles) are built
    g.synchronize()
    launches0 = capi.launch_count()
    g.invoke()                                               # eager again: exactly one step's kernels
    g.synchronize()
    per_step_launches = capi.launch_count() - launches0
    g.enable_cuda_graph(True)
    g.invoke()                                               # the eager pass before the capture
    g.synchronize()
    W = max(args.warmup, 3)
    for _ in range(W):                                       # capture happens on the first
        g.invoke()
    sampler = ClockSampler(D.local)
    if D.rank == 0:
        sampler.start()
    total_ms, w0, w1 = D.timed(g.invoke, args.steps, sync, gs)
    clocks = sampler.stop(w0, w1) if D.rank == 0 else None

    # ---- (2) per-node device times: eager pass with CUDA events on the graph's stream
    g.enable_cuda_graph(False)
    g.enable_profiling(True)
    g.invoke(); g.synchronize(); g.reset_profile()
    prof_ms, _, _ = D.timed(g.invoke, args.steps, sync, gs)
    node_ms = g.node_times_ms()
    g.enable_profiling(False)
    if os.environ.get("LCE_BENCH_VERBOSE") and D.rank == 0:
        for i in range(g.num_nodes()):
            ins, outs = g.node_io(i)
            print(f"node {i:3d} {g.node_name(i):28s} {node_ms[i] / args.steps:8.4f} ms  "
                  f"{g.shape(ins[0])} -> {g.shape(outs[0])}", file=sys.stderr)

    # ---- (3) e2e: pinned host input -> H2D -> graph -> D2H of the result, every step,
    #      double-buffered so the copy of step i+1 overlaps the compute of step i
    e2e_ms = None
    if not args.no_e2e:
        g.enable_cuda_graph(True)
        g.invoke(); g.invoke(); g.synchronize()
        copy_stream = torch.cuda.Stream()
        staging = [torch.empty(host_in.shape, dtype=torch.float32, device=D.dev) for _ in range(2)]
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        state = {"i": 0}

        def issue_copy(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i & 1])
                staging[i & 1].copy_(host_in, non_blocking=True)
                copied[i & 1].record(copy_stream)

        for ev in consumed:
            ev.record(gs)
        issue_copy(0)

        def step_e2e():
            i = state["i"]
            issue_copy(i + 1)                                # prefetch the next step's input
            gs.wait_event(copied[i & 1])
            g.write_ptr(t_in, staging[i & 1].data_ptr(), in_bytes)   # D2D into the arena
            consumed[i & 1].record(gs)
            g.invoke()
            g.read_ptr_async(t_out, host_out.data_ptr(), out_bytes)
            state["i"] = i + 1

        def sync_all():
            g.synchronize()
            copy_stream.synchronize()

        for _ in range(2):
            step_e2e()
        e2e_ms, _, _ = D.timed(step_e2e, args.steps, sync_all, gs)

    # ---- roofline bookkeeping for the dominant kernel (LceBconv2d) ------------------
    conv_ms = conv_bytes = conv_words = 0
    n_conv = 0
    by_op = {}
    for i in range(g.num_nodes()):
        name = g.node_name(i)
        by_op[name] = by_op.get(name, 0.0) + node_ms[i]
        if not name.startswith("LceBconv2d"):
            continue
        ins, outs = g.node_io(i)
        conv_ms += node_ms[i]
        conv_bytes += bconv_alg_bytes(g.shape(ins[0]), g.shape(ins[1]), g.shape(outs[0]))
        if "+ADD" in name:            # fused shortcut read (+ packed signs written)
            conv_bytes += int(np.prod(g.shape(outs[0]))) * 4
        if "+LceQuantize" in name:
            conv_bytes += int(np.prod(g.shape(outs[0]))) // 8
        conv_words += bconv_word_ops(g.shape(outs[0]), g.shape(ins[1]))
        n_conv += 1
    K = args.steps
    imma_default = os.environ.get("LCE_B200_BCONV_IMMA", "1") != "0"
    xor = None
    if imma_default and D.world == 1 and not os.environ.get("LCE_BENCH_NO_XOR_PASS"):
        # same graph with every LceBconv2d plan on the XOR + POPC kernel (north_star's inner
        # product), timed the same way: CUDA-graph replay for the value, eager pass for the kernel
        os.environ["LCE_B200_BCONV_IMMA"] = "0"
        try:
            g2 = H.HostGraph.from_tflite(model_bytes, device_arena=True)
            if not os.environ.get("LCE_NO_FUSION"):
                g2.fuse_all()
            g2.resize_input(g2.inputs()[0], (B, 224, 224, 3))
            g2.allocate_tensors()
            gs2 = torch.cuda.ExternalStream(g2.stream())
            g2.write_ptr(g2.inputs()[0], host_in.data_ptr(), in_bytes)
            g2.synchronize()
            g2.enable_cuda_graph(True)
            for _ in range(max(args.warmup, 3) + 1):
                g2.invoke()
            g2.synchronize()
            x_ms, _, _ = D.timed(g2.invoke, args.steps, g2.synchronize, gs2)
            g2.enable_cuda_graph(False)
            g2.enable_profiling(True)
            g2.invoke(); g2.synchronize(); g2.reset_profile()
            x_prof, _, _ = D.timed(g2.invoke, args.steps, g2.synchronize, gs2)
            x_node = g2.node_times_ms()
            x_conv = sum(x_node[i] for i in range(g2.num_nodes())
                         if g2.node_name(i).startswith("LceBconv2d"))
            xor = {"total_ms": x_ms, "conv_s_per_step": x_conv * 1e-3 / K,
                   "conv_share": x_conv / x_prof if x_prof else None}
            g2.close()
        finally:
            os.environ.pop("LCE_B200_BCONV_IMMA", None)
    return {"total_ms": total_ms, "e2e_ms": e2e_ms, "clocks": clocks, "xor": xor,
            "imma": imma_default,
            "launches": per_step_launches * K, "conv_s_per_step": conv_ms * 1e-3 / K,
            "conv_bytes": conv_bytes, "conv_words": conv_words, "n_conv": n_conv * K,
            "conv_share": conv_ms / prof_ms if prof_ms else None,
            "eager_ms_per_step": prof_ms / K, "in_bytes": in_bytes, "out_bytes": out_bytes,
            "by_op_ms_per_step": {k: round(v / K, 4) for k, v in
                                  sorted(by_op.items(), key=lambda kv: -kv[1])},
            "arena_bytes": g.arena_bytes(), "model_bytes": len(model_bytes),
            "fused_nodes_removed": fused, "graph_nodes": g.num_nodes(),
            "timing_note": "value: CUDA-graph replay; roofline: separate eager pass of the same K "
                           "steps with CUDA events around every node on the graph's stream"}


def stack_workload(args, D):
    torch = D.torch
    from compute_engine_b200 import capi, parallel
    B = args.batch
    layers = make_stack_weights(0) if D.rank == 0 else None
    dev_layers, _ = parallel.broadcast_model(layers, D.dev)    # the only collective
    shapes = [(c, hw) for (hw, c) in STAGES for _ in range(LAYERS_PER_STAGE)]
    plans = []
    for (c, hw), lay in zip(shapes, dev_layers):
        d = capi.BconvDesc(B, hw, hw, c, 3, 3, c, 1, 1, 1, 1, 1, capi.PADDING_SAME, 1,
                           capi.ACT_RELU, capi.OUT_FLOAT, 1.0, 0)
        plans.append(capi.BConv2d(d, lay["filter"], lay["mul"], lay["bias"]))
    host_in = [torch.from_numpy(x).pin_memory() for x in make_stage_inputs(B, 100 + D.rank)]
    dev_in = [x.to(D.dev, non_blocking=True) for x in host_in]
    bufs = [[torch.empty((B, hw, hw, c), device=D.dev) for _ in range(2)] for (hw, c) in STAGES]
    packed = [torch.empty((B, hw, hw, c // 32), dtype=torch.int32, device=D.dev) for (hw, c) in STAGES]
    host_out = [torch.empty((B, hw, hw, c), dtype=torch.float32).pin_memory() for (hw, c) in STAGES]
    ev_pairs = []

    def step(record=False):
        li, outs = 0, []
        for s in range(len(STAGES)):
            x = dev_in[s]
            for k in range(LAYERS_PER_STAGE):
                capi.quantize(x, out=packed[s])
                y = bufs[s][k & 1]
                if record:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); plans[li](packed[s], out=y); e1.record()
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
        for h, o in zip(host_out, step()):
            h.copy_(o, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(D.local)
    if D.rank == 0:
        sampler.start()
    l0 = capi.launch_count()
    total_ms, w0, w1 = D.timed(lambda: step(True), args.steps, torch.cuda.synchronize)
    launches = capi.launch_count() - l0
    clocks = sampler.stop(w0, w1) if D.rank == 0 else None
    torch.cuda.synchronize()
    conv_ms = sum(a.elapsed_time(b) for a, b in ev_pairs)
    e2e_ms = None
    if not args.no_e2e:
        step_e2e()
        e2e_ms, _, _ = D.timed(step_e2e, args.steps, torch.cuda.synchronize)
    conv_bytes = sum(bconv_alg_bytes((B, hw, hw, c // 32), (c, 3, 3, c // 32), (B, hw, hw, c))
                     for hw, c in STAGES) * LAYERS_PER_STAGE
    conv_words = sum(bconv_word_ops((B, hw, hw, c), (c, 3, 3, c // 32))
                     for hw, c in STAGES) * LAYERS_PER_STAGE
    return {"total_ms": total_ms, "e2e_ms": e2e_ms, "clocks": clocks, "launches": launches,
            "conv_s_per_step": conv_ms * 1e-3 / args.steps, "conv_bytes": conv_bytes,
            "conv_words": conv_words, "n_conv": len(ev_pairs), "conv_share": conv_ms / total_ms,
            "in_bytes": sum(x.numel() * 4 for x in host_in),
            "out_bytes": sum(x.numel() * 4 for x in host_out),
            "imma": os.environ.get("LCE_B200_BCONV_IMMA", "1") != "0", "xor": None,
            "timing_note": "CUDA events around every LceBconv2d launch inside the timed region"}


def bgemm_sweep(args, D):
    """BASELINE.json configs[4]: M,N in {256..4096}, K_bits in {256..8192}; rows of A sharded
    over the ranks (W replicated). One JSON object per point on stderr."""
    torch = D.torch
    from compute_engine_b200 import capi
    hbm_peak, _, sm_max = measured_peaks()
    g = torch.Generator(device=D.dev).manual_seed(D.rank)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=D.dev)
    rows = []
    for M in (256, 512, 1024, 2048, 4096):
        for N in (256, 1024, 4096):
            for Kb in (256, 1024, 2048, 8192):
                Kw, Ml = Kb // 32, M // D.world
                A = torch.randint(-2**31, 2**31 - 1, (Ml, Kw), device=D.dev, generator=g,
                                  dtype=torch.int64).to(torch.int32)
                Wt = torch.randint(-2**31, 2**31 - 1, (N, Kw), device=D.dev, generator=g,
                                   dtype=torch.int64).to(torch.int32)
                gemm = capi.BGemm(Wt)
                out = torch.empty((Ml, N), dtype=torch.int32, device=D.dev)
                for _ in range(3):
                    gemm(A, out)
                ts = []
                for _ in range(7):
                    flush.zero_()                       # L2 flush between timed launches
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); gemm(A, out); e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = D.max_ms(float(np.median(ts)))
                alg = (M + N * D.world) * Kw * 4 + M * N * 4
                rows.append({"M": M, "N": N, "K_bits": Kb, "n_gpus": D.world, "ms": round(ms, 5),
                             "binary_TOPS": round(2 * M * N * Kb / ms / 1e9, 2),
                             "alg_GBps": round(alg / ms / 1e6, 1),
                             "hbm_frac": round(alg / ms / 1e6 / (hbm_peak * D.world), 4),
                             "popc_frac": round(M * N * Kw / ms * 1e3 /
                                                (148 * 16 * sm_max * 1e6 * D.world), 3)})
                gemm.close()
                if D.rank == 0:
                    print(json.dumps(rows[-1]), file=sys.stderr)
    return rows


def main_b200(args):
    D = Dist()
    from compute_engine_b200 import capi
    capi.lib()
    hbm_peak, peak_src, sm_max = measured_peaks()
    bf16_peak = measured_bf16_peak()
    popc_peak = 148 * 16 * sm_max * 1e6   # 15.98 POPC/clk/SM measured (profiles/r01_microbench_pipes.jsonl)
    if args.workload == "bgemm_sweep":
        rows = bgemm_sweep(args, D)
        if D.rank == 0:
            best = max(rows, key=lambda r: r["binary_TOPS"])
            emit({"metric": METRIC["bgemm_sweep"], "value": best["binary_TOPS"],
                              "unit": "binary TOPS", "n_gpus": D.world, "steps": 7, "warmup": 3,
                              "ms_per_step": best["ms"], "higher_is_better": True,
                              "scaling": "strong", "vs_baseline": None,
                              "dtype": "u8 x s8 -> s32 on the int8 tensor pipe (== xor-popcount, bit-exact)"
                              if os.environ.get("LCE_B200_BCONV_IMMA", "1") != "0" else "u32 xor-popcount",
                              "data": "synthetic",
                              "config": {"workload": "bgemm_sweep", "best_point": best,
                                         "points": len(rows),
                                         "l2": "192 MiB flush between launches"}})
        if D.world > 1:
            D.dist.destroy_process_group()
        return
    r = (stack_workload if args.workload == "bconv_stack" else graph_workload)(args, D)
    launches = int(D.sum(r["launches"]))
    if D.rank == 0:
        B, K = args.batch, args.steps
        ms_per_step = r["total_ms"] / K
        achieved = r["conv_bytes"] / r["conv_s_per_step"] / 1e9
        line = {
            "metric": METRIC[args.workload], "value": B * D.world / (ms_per_step * 1e-3),
            "unit": "images/s", "n_gpus": D.world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": ("u8 x s8 -> s32 (binary convs on the int8 tensor pipe, bit-exact with "
                      "xor-popcount) + f32 (builtins, epilogue)") if r.get("imma") else
                     "u32 xor-popcount (binary convs) + f32 (builtins, epilogue)",
            "data": "synthetic",
            "config": {"workload": args.workload, "batch_per_gpu": B, "global_batch": B * D.world,
                       "model": f"{args.workload}: synthesised .tflite (random weights, seed 0)",
                       "parallelism": f"dp{D.world}: batch-sharded, one NCCL broadcast of the "
                                      "model at load, no collective per step",
                       "l2": "activations per step exceed the 126 MB L2 (largest tensor 205 MB)"},
            "clocks": r["clocks"],
            "gpu_launches": launches,
        }
        n_launch = max(r["n_conv"] // K, 1)

        def xor_roofline(conv_s, share, traffic):
            ach = r["conv_bytes"] / conv_s / 1e9
            return {"kernel": "lce::bconv_kernel (LceBconv2d, XOR + POPC carry-save tree)",
                    "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "peak_source": peak_src,
                    "traffic": traffic,
                    "traffic_unit": "bytes per launch (mean of the 16 launches of one step, "
                                    "ncu --set full, profiles/r01_ncu_bconv_v8_summary.csv)",
                    "alg_bytes_per_launch": r["conv_bytes"] / n_launch,
                    "launches_timed": r["n_conv"],
                    "avg_launch_ms": conv_s * 1e3 / n_launch,
                    "share_of_step": share, "timing": r["timing_note"],
                    "int_pipe": {"achieved_word_ops_per_s": r["conv_words"] / conv_s,
                                 "naive_popc_peak_per_s": popc_peak,
                                 "frac": r["conv_words"] / conv_s / popc_peak,
                                 "note": "XOR+POPC is POPC-pipe bound (16/clk/SM measured); "
                                         ">1.0 = gain of the carry-save adder tree"}}

        traffic = ncu_traffic_per_launch() if args.workload == "quicknet" else None
        if r.get("imma"):
            # dominant kernel = the int8 tensor-pipe variant of the same inner product
            imma_peak = 148 * 2046.5 * sm_max * 1e6 * 2 / 1e12   # TOP/s, profiles/r01_microbench_mma.jsonl
            tops = 2 * 32 * r["conv_words"] / r["conv_s_per_step"] / 1e12
            line["roofline"] = {
                "kernel": "lce::bconv_imma_kernel (LceBconv2d, u8 x s8 mma.sync m16n8k32 on bitpacked "
                          "operands, bit-exact)",
                "bound": "tensor", "achieved": tops, "peak": imma_peak, "unit": "TOP/s (int8)",
                "frac": tops / imma_peak,
                "peak_source": "measured here: legacy int8 mma.sync 2046.5 MAC/clk/SM x 148 SMs x "
                               "SM clock (tools/microbench_mma.cu, profiles/r01_microbench_mma.jsonl); "
                               "MEASURED_PEAKS.json's bf16 figure is the tcgen05 path and is not a bound "
                               "for mma.sync int8",
                "traffic": ncu_traffic_per_launch("r01_ncu_bconv_imma_summary.csv")
                if args.workload == "quicknet" else None,
                "traffic_unit": "bytes per launch (ncu --set full, "
                                "profiles/r01_ncu_bconv_imma_summary.csv)",
                "alg_bytes_per_launch": r["conv_bytes"] / n_launch,
                "hbm": {"achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak},
                "vs_measured_bf16_dense": {"peak": bf16_peak, "unit": "TFLOP/s (MEASURED_PEAKS.json, tcgen05 "
                                           "path; shown for scale only)",
                                           "frac": (tops / bf16_peak) if bf16_peak else None},
                "launches_timed": r["n_conv"],
                "avg_launch_ms": r["conv_s_per_step"] * 1e3 / n_launch,
                "share_of_step": r["conv_share"], "timing": r["timing_note"]}
            if r.get("xor"):
                x = r["xor"]
                line["xor_popc_path"] = {
                    "note": "same graph, same run, LCE_B200_BCONV_IMMA=0: every LceBconv2d on the "
                            "XOR + POPC kernel north_star describes",
                    "value": B * D.world / (x["total_ms"] / K * 1e-3), "unit": "images/s",
                    "ms_per_step": x["total_ms"] / K,
                    "roofline": xor_roofline(x["conv_s_per_step"], x["conv_share"], traffic)}
        else:
            line["roofline"] = xor_roofline(r["conv_s_per_step"], r["conv_share"], traffic)
        if r.get("e2e_ms") is not None:
            line["e2e"] = {"value": B * D.world / (r["e2e_ms"] / K * 1e-3), "unit": "images/s",
                           "h2d_bytes_per_step": r["in_bytes"],
                           "d2h_bytes_per_step": r["out_bytes"], "ms_per_step": r["e2e_ms"] / K,
                           "note": "pinned host input -> H2D -> graph -> D2H result every step; "
                                   "the copy of step i+1 overlaps the compute of step i"}
        for k in ("by_op_ms_per_step", "eager_ms_per_step", "arena_bytes", "model_bytes",
                  "fused_nodes_removed", "graph_nodes"):
            if k in r:
                line["config"][k] = r[k]
        if not args.no_cpu_baseline and D.world == 1:
            line["cpu_baseline"] = cpu_baseline_subprocess(args.workload, B)
        emit(line)
    if D.world > 1:
        D.dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
