"""Build libgfb.so (nvcc, sm_100a only) in-tree.  One translation unit per FFT length, compiled in
parallel.  No CPU fallback is built here: the product requires the CUDA library."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libgfb.so")
SIZES = [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + \
        [os.path.join(HERE, "..", "include", "gfb.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout[-4000:]))
    return r.stdout


def build(force=False, verbose=False, nvcc="nvcc"):
    os.makedirs(OBJ, exist_ok=True)
    deps = _sources()
    jobs = []
    for n in SIZES:
        obj = os.path.join(OBJ, "launchers_%d.o" % n)
        if force or _stale(obj, deps):
            jobs.append(([nvcc] + NVCC_FLAGS + ["-DGFB_N=%d" % n, "-c", os.path.join(CSRC, "launchers.cu"), "-o", obj],
                         obj[:-2] + ".log"))
    obj = os.path.join(OBJ, "gfb.o")
    if force or _stale(obj, deps):
        jobs.append(([nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, "gfb.cu"), "-o", obj], obj[:-2] + ".log"))
    obj = os.path.join(OBJ, "cols_dft.o")
    if force or _stale(obj, deps):
        jobs.append(([nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, "cols_dft.cu"), "-o", obj], obj[:-2] + ".log"))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(lambda j: _run(*j), jobs))
    objs = [os.path.join(OBJ, "launchers_%d.o" % n) for n in SIZES] + [os.path.join(OBJ, "gfb.o"), os.path.join(OBJ, "cols_dft.o")]
    if force or jobs or _stale(LIB, objs):
        _run([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"], os.path.join(OBJ, "link.log"))
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
