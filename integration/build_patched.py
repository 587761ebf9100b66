#!/usr/bin/env python
"""Builds oracle/_ref/PHI_gpu: the reference CLI with its front end (ILP_index.cpp:543-743) replaced by a call into
libphi_gpu_index.so through integration/phi_adapter.hpp.  TEST INFRASTRUCTURE: it exists to prove the drop-in
(tests/test_gpu_dropin.py compares its model dump with the unmodified reference's, byte for byte).

The patched translation unit is generated in a temporary directory from /root/reference/src/ILP_index.cpp and deleted
again; no reference source is copied into the repo.  Everything else is compiled from the reference sources in place,
against the same recording Gurobi stand-in as oracle/_ref/PHI_ref.
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PHI_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src")
OUT = os.path.join(ROOT, "oracle", "_ref")
START = "std::vector<int32_t> hap_sizes(num_walks);"
END = "(float)retained_kmers/(float)count_sp_r * 100);"


def main():
    lines = open(os.path.join(SRC, "ILP_index.cpp")).read().split("\n")
    s = next(i for i, l in enumerate(lines) if START in l)
    e = next(i for i, l in enumerate(lines) if END in l)
    assert (s + 1, e + 1) == (543, 743), f"seam moved: {s + 1}-{e + 1}"
    inc = next(i for i, l in enumerate(lines) if '#include "ILP_index.h"' in l)
    seam = open(os.path.join(ROOT, "integration", "seam.inc")).read().rstrip("\n").split("\n")
    patched = lines[:inc + 1] + ['#include "phi_adapter.hpp"'] + lines[inc + 1:s] + seam + lines[e + 1:]
    os.makedirs(OUT, exist_ok=True)
    flags = ["-std=c++11", "-fopenmp", "-pthread", "-O3", "-march=x86-64-v2", "-mtune=generic", "-w",
             "-I", os.path.join(ROOT, "oracle", "ref_build", "stub"), "-I", SRC, "-I", os.path.join(ROOT, "include"),
             "-I", os.path.join(ROOT, "integration")]
    with tempfile.TemporaryDirectory() as tmp:
        cpp = os.path.join(tmp, "ILP_index_gpu.cpp")
        open(cpp, "w").write("\n".join(patched))
        obj = os.path.join(tmp, "ILP_index_gpu.o")
        subprocess.check_call(["g++"] + flags + ["-c", cpp, "-o", obj])
        others = [os.path.join(OUT, "obj", f + ".o") for f in ("gfa-io", "gfa-base", "options", "kalloc", "misc", "sys", "MurmurHash3", "main")]
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
        libdir = os.path.join(ROOT, "phi_b200")
        subprocess.check_call(["g++"] + flags + [obj] + others + ["-o", os.path.join(OUT, "PHI_gpu"), "-L", libdir, "-lphi_gpu_index",
                               "-Wl,-rpath,$ORIGIN/../../phi_b200", "-lm", "-lz", "-lpthread", "-ldl"])
    print(os.path.join(OUT, "PHI_gpu"))


if __name__ == "__main__":
    sys.exit(main())
