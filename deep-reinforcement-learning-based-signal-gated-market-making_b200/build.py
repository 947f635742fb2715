"""Build recipe of libsgmm_b200.so (nvcc, sm_100a only, in-tree).

    python -m sgmm_b200.build            # or: __graft_entry__.build()

The library is a plain C-ABI shared object (include/sgmm.h); cudart is linked statically so the
only run-time dependency is the NVIDIA driver.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsgmm_b200.so")
SOURCES = ["sgmm_capi.cu", "sgmm_rollout.cu", "sgmm_spec256.cu", "sgmm_tc32.cu", "sgmm_ga.cu", "sgmm_peak.cu", "sgmm_prep.cu", "sgmm_account.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--fmad=false",                       # no implicit a*b+c contraction: every FMA is written as one
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-Xptxas", "-v",
    "-cudart", "static",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libsgmm_b200.so cannot be built")
    return p


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "sgmm.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    extra = os.environ.get("SGMM_EXTRA_NVCC_FLAGS", "").split()      # debug builds, e.g. -DSGMM_TC32_WATCHDOG / -DSGMM_SPEC256_TRACE
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + ["-shared", "-o", LIB] + srcs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "csrc", "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed (exit {r.returncode}); see {log}")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
