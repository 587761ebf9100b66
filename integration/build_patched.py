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
# second variant (PHI_gpu_model): additionally the k-mer constraint block of the model construction, :782-880
BLOCK_START = "        if (is_ilp)"
BLOCK_END = "// print count_sp_r_ilp/count_sp_r * 100% kmer matches are in ilp"
# and the "optimized expanded graph" part of the default branch, :1201-1406
# the naive expanded graph (-N1), :942-1154: from the first "// w/o recombination" to the line before the first "// clear vars"
NAIVE_START = "            // w/o recombination"
NAIVE_END = "            // clear vars"
GRAPH_START = "            std::map<std::string, std::vector<std::string>> new_adj;"
GRAPH_END = "            in_nodes_new.clear();"


def main():
    lines = open(os.path.join(SRC, "ILP_index.cpp")).read().split("\n")
    s = next(i for i, l in enumerate(lines) if START in l)
    e = next(i for i, l in enumerate(lines) if END in l)
    assert (s + 1, e + 1) == (543, 743), f"seam moved: {s + 1}-{e + 1}"
    inc = next(i for i, l in enumerate(lines) if '#include "ILP_index.h"' in l)
    bs = next(i for i, l in enumerate(lines) if l.rstrip() == BLOCK_START)
    be = next(i for i, l in enumerate(lines) if BLOCK_END in l) - 2                     # the closing brace of the else branch
    assert (bs + 1, be + 1) == (782, 880) and lines[be].strip() == "}", f"model block moved: {bs + 1}-{be + 1}"
    ns = next(i for i, l in enumerate(lines) if l.rstrip() == NAIVE_START)
    ne = next(i for i, l in enumerate(lines) if l.rstrip() == NAIVE_END) - 1
    assert (ns + 1, ne + 1) == (942, 1155) and lines[ne].strip() == "", f"naive block moved: {ns + 1}-{ne + 1}"
    gs = next(i for i, l in enumerate(lines) if l.rstrip() == GRAPH_START)
    ge = next(i for i, l in enumerate(lines) if l.rstrip() == GRAPH_END)
    assert (gs + 1, ge + 1) == (1201, 1406), f"expanded-graph block moved: {gs + 1}-{ge + 1}"

    def read(name):
        return open(os.path.join(ROOT, "integration", name)).read().rstrip("\n").split("\n")
    variants = {
        # the drop-in proper: only the front end is replaced
        "PHI_gpu": (lines[:inc + 1] + ['#include "phi_adapter.hpp"'] + lines[inc + 1:s] + read("seam.inc") + lines[e + 1:], []),
        # + the k-mer constraint block built straight from the result; test hook compiled in
        "PHI_gpu_model": (lines[:inc + 1] + ['#include "phi_model.hpp"'] + lines[inc + 1:s] + read("seam_model.inc") + lines[e + 1:bs]
                          + read("model_block.inc") + lines[be + 1:ns] + read("model_naive.inc") + lines[ne:gs] + read("model_graph.inc") + lines[ge + 1:],
                          ["-DPHI_ADAPTER_TESTHOOK"]),
    }
    os.makedirs(OUT, exist_ok=True)
    flags = ["-std=c++11", "-fopenmp", "-pthread", "-O3", "-march=x86-64-v2", "-mtune=generic", "-w",
             "-I", os.path.join(ROOT, "oracle", "ref_build", "stub"), "-I", SRC, "-I", os.path.join(ROOT, "include"),
             "-I", os.path.join(ROOT, "integration")]
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    others = [os.path.join(OUT, "obj", f + ".o") for f in ("gfa-io", "gfa-base", "options", "kalloc", "misc", "sys", "MurmurHash3", "main")]
    libdir = os.path.join(ROOT, "phi_b200")
    with tempfile.TemporaryDirectory() as tmp:
        for name, (patched, extra) in variants.items():
            cpp = os.path.join(tmp, name + ".cpp")
            open(cpp, "w").write("\n".join(patched))
            obj = os.path.join(tmp, name + ".o")
            subprocess.check_call(["g++"] + flags + extra + ["-c", cpp, "-o", obj])
            subprocess.check_call(["g++"] + flags + [obj] + others + ["-o", os.path.join(OUT, name), "-L", libdir, "-lphi_gpu_index",
                                   "-Wl,-rpath,$ORIGIN/../../phi_b200", "-lm", "-lz", "-lpthread", "-ldl"])
            print(os.path.join(OUT, name))


if __name__ == "__main__":
    sys.exit(main())
