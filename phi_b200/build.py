"""Builds phi_b200/libphi_gpu_index.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

No torch.utils.cpp_extension: the library has no torch types in it; it is loaded with ctypes.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libphi_gpu_index.so")
SOURCES = ["read_sketch.cu", "walk_sketch.cu", "chunks.cu", "primitives.cu", "filter.cu", "groups.cu", "phi_gpu_index.cu", "shard.cu", "host_io.cpp", "merge.cpp"]
HEADERS = ["kernels.h", "device_common.cuh", "sketch_tile.cuh", "sketch_common.cuh", "result_box.h", "fast_inflate.h", os.path.join(ROOT, "include", "phi_gpu_index.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o").replace(".cpp", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs + [os.path.abspath(__file__)]):
            cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append(f"== {s}\n{out}")
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {s}")
    if procs or force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "--cudart", "static", "-lz"]
        subprocess.check_call(cmd)
    if log:
        with open(os.path.join(objdir, "ptxas.log"), "w") as f:
            f.write("\n".join(log))
        if verbose:
            print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
