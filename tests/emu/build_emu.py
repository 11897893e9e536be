"""TEST INFRASTRUCTURE: compile the product's kernel sources with g++ through cuda_emu.h into
tests/emu/libgfb_emu.so so that kernel logic can be checked against the oracle without a GPU.
Never used by the product package."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "geophysicalflows.jl_b200", "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libgfb_emu.so")
SIZES = [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]
FLAGS = ["-std=c++17", "-O1", "-fPIC", "-DGFB_EMULATE", "-x", "c++", "-I", HERE, "-I", CSRC, "-w"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("emu build failed: %s\n%s" % (" ".join(cmd), r.stdout[-6000:]))


def build(force=False):
    os.makedirs(OBJ, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps += [os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "gfb.h")]
    jobs = []
    for n in SIZES:
        obj = os.path.join(OBJ, "launchers_%d.o" % n)
        if force or _stale(obj, deps):
            jobs.append(["g++"] + FLAGS + ["-DGFB_N=%d" % n, "-c", os.path.join(CSRC, "launchers.cu"), "-o", obj])
    obj = os.path.join(OBJ, "gfb.o")
    if force or _stale(obj, deps):
        jobs.append(["g++"] + FLAGS + ["-DGFB_EMU_IMPLEMENTATION", "-c", os.path.join(CSRC, "gfb.cu"), "-o", obj])
    obj = os.path.join(OBJ, "cols_dft.o")
    if force or _stale(obj, deps):
        jobs.append(["g++"] + FLAGS + ["-c", os.path.join(CSRC, "cols_dft.cu"), "-o", obj])
    if jobs:
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            list(ex.map(_run, jobs))
    objs = [os.path.join(OBJ, "launchers_%d.o" % n) for n in SIZES] + [os.path.join(OBJ, "gfb.o"), os.path.join(OBJ, "cols_dft.o")]
    if force or jobs or _stale(LIB, objs):
        _run(["g++", "-shared", "-o", LIB] + objs + ["-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build())
