"""Development: build build/liblce_b200_prof.so = the CUDA library with the per-role cycle
counters compiled in (-DLCE_TC_PROF=1); use with LCE_B200_LIB=build/liblce_b200_prof.so
LCE_B200_TC_PROF=1."""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from compute_engine_b200 import build as B  # noqa: E402

os.makedirs(os.path.join(REPO, "build"), exist_ok=True)
cus = sorted(os.path.join(B.CSRC, f) for f in os.listdir(B.CSRC) if f.endswith(".cu"))
out = os.path.join(REPO, "build", "liblce_b200_prof.so")
subprocess.run([B._nvcc(), "-DLCE_TC_PROF=1", *B.NVCC_FLAGS, "-I", B.INC, *cus, "-o", out], check=True)
print(out)
