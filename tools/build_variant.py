"""Development: build build/liblce_b200_<tag>.so with extra nvcc defines for A/B runs
(LCE_B200_LIB=build/liblce_b200_<tag>.so). Usage: build_variant.py tag -DNAME=VALUE ..."""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from compute_engine_b200 import build as B  # noqa: E402

tag, defs = sys.argv[1], sys.argv[2:]
os.makedirs(os.path.join(REPO, "build"), exist_ok=True)
cus = sorted(os.path.join(B.CSRC, f) for f in os.listdir(B.CSRC) if f.endswith(".cu"))
out = os.path.join(REPO, "build", f"liblce_b200_{tag}.so")
subprocess.run([B._nvcc(), *defs, *B.NVCC_FLAGS, "-I", B.INC, *cus, "-o", out], check=True)
print(out)
