"""In-tree build of libcofusion_b200.so (hand-written CUDA for sm_100a + the C ABI).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the gpurun snapshot.
Per-file flags: files listed in NO_FMAD are compiled with -fmad=false so that the values feeding
integer decisions (index maps, association, cleaning, labels) are bit-identical to the CPU oracle.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libcofusion_b200.so")
NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr"]
NO_FMAD = {"image_kernels.cu", "surfel_kernels.cu", "segment_kernels.cu"}


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "cofusion_b200.h"))
    headers.append(os.path.abspath(__file__))
    objs = []
    for src in _sources():
        o = os.path.join(OBJ, src[:-3] + ".o")
        objs.append(o)
        if force or _newer(o, [os.path.join(CSRC, src)] + headers):
            cmd = [NVCC] + ARCH + COMMON + (["-fmad=false"] if src in NO_FMAD else []) + \
                  ["-Xptxas", "-v"] * bool(verbose) + ["-c", os.path.join(CSRC, src), "-o", o]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose and r.stderr:
                print(r.stderr, file=sys.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if force or _newer(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
