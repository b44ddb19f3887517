#!/usr/bin/env python
"""Turn the raw evidence a `tools/capture_evidence.sh` run left under gpurun_out/ev/ into the
committed summaries under profiles/ (round 2). Runs here (ncu -i needs no GPU).
Usage: python tools/make_profiles.py [tag]   (tag defaults to "r02")"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EV = os.path.join(REPO, "gpurun_out", "ev")
OUT = os.path.join(REPO, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg"]


def ncu_summary(rep, dst):
    path = os.path.join(EV, rep)
    if not os.path.exists(path):
        print("missing", rep)
        return
    if path.endswith(".csv"):      # exported on the GPU box (the reports are too large to travel)
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    with open(os.path.join(OUT, dst), "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            if len(r) >= len(hdr):
                w.writerow([r[c] for c in cols])
    print("wrote", dst, len(rows) - 2, "launches")


def launch_list(src, dst):
    path = os.path.join(EV, src)
    if not os.path.exists(path):
        print("missing", src)
        return
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
    tot, n = {}, {}
    for r in rows[1:]:
        name = r[ik].split("(")[0]
        us = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
        tot[name] = tot.get(name, 0.0) + us
        n[name] = n.get(name, 0) + 1
    total = sum(tot.values())
    with open(os.path.join(OUT, dst), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share"])
        for k in sorted(tot, key=lambda k: -tot[k]):
            w.writerow([k, n[k], round(tot[k], 1), round(tot[k] / total, 4)])
    shutil.copy(path, os.path.join(OUT, dst.replace("_summary", "")))
    print("wrote", dst, "total", round(total), "us over", sum(n.values()), "launches")


def copy(src, dst):
    path = os.path.join(EV, src)
    if os.path.exists(path):
        shutil.copy(path, os.path.join(OUT, dst))
        print("copied", dst)
    else:
        print("missing", src)


def main():
    ncu_summary("ncu_bconv_tc_raw.csv", f"{TAG}_ncu_bconv_tc_summary.csv")
    ncu_summary("ncu_glue_raw.csv", f"{TAG}_ncu_glue_summary.csv")
    ncu_summary("ncu_aux_raw.csv", f"{TAG}_ncu_aux_kernels_summary.csv")
    ncu_summary("ncu_stem7_raw.csv", f"{TAG}_ncu_stem7_summary.csv")
    launch_list("ncu_launches.csv", f"{TAG}_ncu_launches_summary.csv")
    for src, dst in (("bench_default.json", f"{TAG}_bench_quicknet_b256_final.json"),
                     ("bench_reference.json", f"{TAG}_bench_reference_arm.json"),
                     ("bench_1000steps.json", f"{TAG}_bench_quicknet_b256_1000steps.json"),
                     ("bgemm_sweep.jsonl", f"{TAG}_bgemm_sweep_1gpu.jsonl"),
                     ("bench_bgemm_sweep.json", f"{TAG}_bench_bgemm_sweep.json"),
                     ("tc_check.log", f"{TAG}_tc_check_final.txt"), ("tc_prof.log", f"{TAG}_tc_roles_final.txt"),
                     ("pw_check.log", f"{TAG}_pw_check.txt"), ("pw_prof.log", f"{TAG}_pw_roles.txt"),
                     ("stem_check.log", f"{TAG}_stem_check.txt"), ("tc_probe.log", f"{TAG}_tc_probe.txt"),
                     ("pytest_gpu.log", f"{TAG}_pytest_gpu.txt"), ("sanitizer_builtins.log", f"{TAG}_compute_sanitizer_builtins.txt"),
                     ("sanitizer_tc.log", f"{TAG}_compute_sanitizer_tc.txt"), ("nvidia_smi.txt", f"{TAG}_nvidia_smi.txt")):
        copy(src, dst)
    # per-node table from the verbose bench
    p = os.path.join(EV, "bench_nodes.err")
    if os.path.exists(p):
        with open(os.path.join(OUT, f"{TAG}_bench_nodes.txt"), "w") as f:
            f.writelines(ln for ln in open(p) if ln.startswith("node"))
        print("wrote nodes")
    # SASS evidence: tensor-core / TMA mnemonics per kernel, from the built library
    lib = os.path.join(REPO, "compute_engine_b200", "liblce_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, counts = None, {}
    pats = ("UTCIMMA", "UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "LDGSTS", "IMMA", "POPC")
    for ln in sass.splitlines():
        if "Function :" in ln:
            cur = ln.split("Function :")[1].strip()
            counts[cur] = {}
        elif cur:
            for pat in pats:
                if pat in ln:
                    counts[cur][pat] = counts[cur].get(pat, 0) + 1
    with open(os.path.join(OUT, f"{TAG}_sass_mnemonics.txt"), "w") as f:
        f.write("cuobjdump -sass compute_engine_b200/liblce_b200.so: instruction mnemonics per kernel (sm_100a)\n")
        for k, c in counts.items():
            if any(p in c for p in ("UTCIMMA", "UTCHMMA", "FFMA2", "UTMALDG", "IMMA", "UBLKCP")):
                name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:160]
                f.write(f"{name}\n    {json.dumps(c, sort_keys=True)}\n")
    print("wrote sass mnemonics")
    # and the instructions themselves for the three tcgen05 kernels
    want = ("bconv_tc_kernelILi4ELi0", "pw_tf32_kernel", "stem7_tf32_kernel")
    with open(os.path.join(OUT, f"{TAG}_sass_tcgen05_excerpt.txt"), "w") as f:
        f.write("cuobjdump -sass compute_engine_b200/liblce_b200.so, lines with tcgen05 / TMA / TMEM instructions\n")
        cur = None
        for ln in sass.splitlines():
            if "Function :" in ln:
                cur = ln.split("Function :")[1].strip()
                if any(w in cur for w in want):
                    f.write("\n== " + cur + "\n")
            elif cur and any(w in cur for w in want) and any(t in ln for t in ("UTCIMMA", "UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMACMDFLUSH")):
                f.write(ln.rstrip()[:140] + "\n")
    print("wrote sass excerpt")


if __name__ == "__main__":
    main()
