"""In-tree build of the C-ABI CUDA library (nvcc, sm_100a only).  No torch C++ API is used: the
library exposes plain `extern "C"` symbols (include/seg2eye_b200.h) and is bound with ctypes."""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB = os.path.join(OUT_DIR, "libseg2eye_b200.so")
SOURCES = ["misc.cu", "norm.cu", "conv_simt.cu", "conv_thin.cu", "conv_tc.cu", "tail.cu", "data.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "--extended-lambda"]
# --use_fast_math only where it pays and cannot be seen: the convolution epilogues and the streaming normalisation kernels
# (bf16 outputs).  tanh of the image head, losses, Adam, the spectral iteration, the validation tail and the data layer are
# compiled with IEEE division / sqrt / transcendental functions.
FAST_MATH = {"conv_tc.cu", "conv_simt.cu", "norm.cu"}


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(os.path.dirname(HERE), "include", "seg2eye_b200.h")]
    nvcc = _nvcc()
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OUT_DIR, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + (["--use_fast_math"] if s in FAST_MATH else []) + (["-Xptxas", "-v"] if verbose else [])
                        + ["-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=4) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    if jobs or force or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
