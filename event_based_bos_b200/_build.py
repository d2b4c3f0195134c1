"""Build libebos.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m event_based_bos_b200._build [--force]

The library is a plain shared object (no torch, no pybind): `extern "C"` entry points declared in
include/ebos.h.  cudart is linked statically, so the .so loads on a machine without a GPU (the
entry points then return EBOS_ERR_CUDA).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libebos.so")
SOURCES = ["ebos_ops.cu", "ebos_window.cu", "ebos_costs.cu", "ebos_ingest.cu", "ebos_eklt.cu", "ebos_metrics.cu"]
HEADERS = ["ebos_common.cuh", "ebos_eklt_math.cuh", os.path.join("..", "..", "include", "ebos.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
]


# No per-file flags at present.  (ptxas 12.9 contracts explicit mul.rn.f32x2 + sub.rn.f32x2 into FFMA2 even with
# --fmad=false; the exactness-critical warp therefore uses the scalar __fmul_rn/__fsub_rn forms, see ebos_window.cu.)
PER_FILE_FLAGS = {}
if os.environ.get("EBOS_BUILD_GM_GROUPS"):      # experiment: row groups per CTA of the gradient-magnitude kernel
    NVCC_FLAGS.append("-DEBOS_GM_GROUPS=" + os.environ["EBOS_BUILD_GM_GROUPS"])
if os.environ.get("EBOS_BUILD_EKLT_MINB"):    # experiment: occupancy cap of the EKLT per-pixel kernels (csrc/ebos_eklt.cu)
    NVCC_FLAGS.append("-DEBOS_EKLT_MINB=" + os.environ["EBOS_BUILD_EKLT_MINB"])
if os.environ.get("EBOS_BUILD_GATHER_SELECT"):   # experiment: flow gathers into registers of their own + selects (ebos_window.cu)
    NVCC_FLAGS.append("-DEBOS_GATHER_SELECT")
if os.environ.get("EBOS_BUILD_NO_RED_V2"):       # A/B: scalar reductions only (default: red.global.add.v2.f32 for column pairs, ebos_window.cu)
    NVCC_FLAGS.append("-DEBOS_NO_RED_V2")
if os.environ.get("EBOS_BUILD_OUT"):             # A/B builds: write the library somewhere else (EBOS_LIBRARY selects it at run time)
    LIB_PATH = os.environ["EBOS_BUILD_OUT"]
if os.environ.get("EBOS_BUILD_ABLATION"):   # diagnostics build: EBOS_ABLATE=<mask> then removes kernel components
    NVCC_FLAGS.append("-DEBOS_ABLATION")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    deps.append(os.path.abspath(__file__))   # the source list and the flags live here
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into event_based_bos_b200/libebos.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    build_dir = os.path.join(PKG_DIR, "csrc", "build" + ("_alt" if os.environ.get("EBOS_BUILD_OUT") else ""))
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for s in SOURCES:
        obj = os.path.join(build_dir, s.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *PER_FILE_FLAGS.get(s, []), "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
           "-Xcompiler", "-fPIC", "-o", tmp, *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
