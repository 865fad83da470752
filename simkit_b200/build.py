"""Builds the CUDA library in-tree with nvcc for sm_100a (no JIT cache: the .so travels).

    python -m simkit_b200.build            # build libsimkit_b200.so if stale
    python -m simkit_b200.build --force
"""

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# development switches: SKB_BUILD_TAG=name builds libsimkit_b200_name.so (objects in _obj_name) with the extra
# nvcc flags of SKB_BUILD_FLAGS; the loader picks it up through SKB_LIB_TAG (A/B kernel experiments)
_TAG = os.environ.get("SKB_BUILD_TAG", "")
OBJ = os.path.join(HERE, "csrc", "_obj" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(HERE, "libsimkit_b200" + ("_" + _TAG if _TAG else "") + ".so")
SOURCES = ["capi.cu", "capi_elements.cu", "capi_solver.cu", "capi_reduced.cu", "capi_dist.cu", "capi_nccl.cu", "capi_pcg2.cu", "capi_buffers.cu", "capi_subspace.cu", "capi_cluster.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "--expt-extended-lambda",
    "-Xcompiler", "-fPIC",
    "-Xcudafe", "--diag_suppress=177",
] + os.environ.get("SKB_BUILD_FLAGS", "").split()


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    return nvcc


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "simkit_b200.h"))
    return out


def is_stale():
    if not os.path.exists(LIB):
        return True
    lib_t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > lib_t for p in _deps())


def build(force=False, verbose=False, ptxas_info=False):
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps_t = max(os.path.getmtime(p) for p in _deps() if p.endswith((".cuh", ".h")))

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        spath = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(spath), deps_t):
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", spath, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if verbose or ptxas_info:
        for _, log in results:
            if log:
                print(log)
    # cuSOLVER: dense Cholesky inverse of the coarse system of the two-level PCG preconditioner (csrc/coarse.cuh)
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart", "-lcusolver", "-lcublas", "-ldl", "-Xlinker", "-rpath=" + cuda_lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv)
    print("built", path)
