"""Builds the sm_100a shared library of the engine in-tree: surikatoko_b200/_lib/libsrk_ba.so.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the repository snapshot.
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "_lib")
LIB = os.path.join(LIB_DIR, "libsrk_ba.so")
SOURCES = ["engine.cu", "ba_kernels.cu", "schur_mma.cu", "schur_v3.cu", "prep_kernels.cu", "chol_kernels.cu", "pcg_kernels.cu", "aux_kernels.cu", "ekf_kernels.cu", "frontend.cu", "solve_order.cu", "bundle.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
              "--expt-relaxed-constexpr"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".inc"))]
    headers += [os.path.join(_HERE, "..", "include", "srk", f) for f in os.listdir(os.path.join(_HERE, "..", "include", "srk"))]
    objs = []
    procs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(LIB_DIR, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s\n%s\n" % (s, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-Wno-deprecated-gpu-targets", "-o", LIB] + objs + ["-lcudart", "-ldl"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
