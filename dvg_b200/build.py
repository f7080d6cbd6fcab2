"""Build recipe for libdvg_b200.so (nvcc, sm_100a only, in-tree so it travels with gpurun snapshots)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
# DVG_LIB_TAG=<tag>: developer builds (DVG_TRACE=1, DVG_STEP_*=...) go to _lib_<tag>/ and are loaded from there, so a
# trace build and the product build can travel to the GPU box side by side.
LIB_DIR = os.path.join(PKG, "_lib" + ("_" + os.environ["DVG_LIB_TAG"] if os.environ.get("DVG_LIB_TAG") else ""))
LIB = os.path.join(LIB_DIR, "libdvg_b200.so")
SOURCES = ["capi.cu", "lstm_fp32.cu", "lstm_tc.cu", "lstm_step.cu", "lstm_small.cu", "gp.cu", "gp_big.cu", "gp_tc.cu", "gp_factor.cu", "rollout.cu", "moving_mnist.cu"]
HEADERS = ["common.cuh", "internal.cuh", "ptx.cuh", "gp_trigger.cuh", "gp_rsample.cuh", "tc_common.cuh", os.path.join("..", "..", "include", "dvg_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas", "-v", "-cudart", "static",
]
if os.environ.get("DVG_STEP_NPOLY"):     # developer build: exponentials on the FMA pipe in the step kernel's epilogue
    NVCC_FLAGS.append("-DDVG_STEP_NPOLY=" + os.environ["DVG_STEP_NPOLY"])
if os.environ.get("DVG_STEP_TRIG_EARLY"):   # developer build: 0 = trigger partial sums after the first tile
    NVCC_FLAGS.append("-DDVG_STEP_TRIG_EARLY=" + os.environ["DVG_STEP_TRIG_EARLY"])
if os.environ.get("DVG_STEP_EW"):        # developer build: 8 or 16 epilogue warps in the step kernel
    NVCC_FLAGS.append("-DDVG_STEP_EW=" + os.environ["DVG_STEP_EW"])
for _k in ("DVG_STEP_POLL_BATCH", "DVG_STEP_WFENCE", "DVG_STEP_WARM", "DVG_STEP_X"):      # developer builds: A/B switches of the step kernel
    if os.environ.get(_k):
        NVCC_FLAGS.append("-D%s=%s" % (_k, os.environ[_k]))
if os.environ.get("DVG_TRACE"):          # developer build: per-CTA timestamps in the tensor-core kernel
    NVCC_FLAGS.append("-DDVG_TRACE")


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found; dvg_b200 needs the CUDA toolkit to build its kernels")


def source_hash() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def is_current() -> bool:
    stamp = LIB + ".hash"
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == source_hash()


def build(force: bool = False, verbose: bool = True) -> str:
    """Compile every .cu for sm_100a into dvg_b200/_lib/libdvg_b200.so (object files compiled in parallel)."""
    if not force and is_current():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", LIB, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(os.path.join(LIB_DIR, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(LIB + ".hash", "w") as fh:
        fh.write(source_hash())
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
