#!/usr/bin/env python
"""SURVEY 8(f) rows 1-2 measurement: host ingest of the README fixture (test/MHC_4.gfa.gz + test/CHM13_reads.fq.gz).
Reference: the unmodified CLI's own log stamp "Graph has ..." (gfa_read + ILP_index::read_gfa + kseq read loop, single thread;
oracle/_ref/PHI_ref, killed once the stamp is printed).  Ours: phi_host_graph_load + phi_host_reads_load (phi_b200/csrc/host_io.cpp)
through ctypes, median of 7.  Needs /root/reference (the fixture files live there)."""
import ctypes as C
import json
import os
import re
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phi_b200  # noqa: E402

GFA, FQ = "/root/reference/test/MHC_4.gfa.gz", "/root/reference/test/CHM13_reads.fq.gz"


def ours():
    lib = phi_b200.load_library()
    tg, tr = [], []
    for _ in range(7):
        h = C.c_void_p(); err = C.create_string_buffer(256)
        t0 = time.perf_counter(); rc = lib.phi_host_graph_load(GFA.encode(), C.byref(h), err, 256); t1 = time.perf_counter()
        assert rc == 0, err.value
        lib.phi_host_graph_free(h)
        h = C.c_void_p()
        t2 = time.perf_counter(); rc = lib.phi_host_reads_load(FQ.encode(), C.byref(h), err, 256); t3 = time.perf_counter()
        assert rc == 0, err.value
        lib.phi_host_reads_free(h)
        tg.append(t1 - t0); tr.append(t3 - t2)
    return statistics.median(tg), statistics.median(tr)


def reference():
    vals = []
    for _ in range(5):
        p = subprocess.Popen([os.path.join(ROOT, "oracle", "_ref", "PHI_ref"), "-g", GFA, "-r", FQ, "-o", "/tmp/ingest_ref.fa", "-t", "1"],
                             stderr=subprocess.PIPE, text=True, env=dict(os.environ, PHI_STUB_DUMP="/tmp/ingest_ref.dump"))
        loaded = None
        for line in p.stderr:
            m = re.match(r"\[M::main::([\d.]+)\*", line)
            if m and "Loaded graph" in line:
                loaded = float(m.group(1))
            m = re.match(r"\[M::ILP_function::([\d.]+)\*[\d.]+\] Graph has", line)
            if m:
                vals.append((loaded, float(m.group(1))))
                break
        p.kill(); p.wait()
    return statistics.median(v[0] for v in vals), statistics.median(v[1] for v in vals)


if __name__ == "__main__":
    g, r = ours()
    rl, rt = reference()
    print(json.dumps({"fixture": "test/MHC_4.gfa.gz (3.4 MB gz, 14 MB text, 111,805 S / 151,740 L / 5 W lines) + test/CHM13_reads.fq.gz (16,401 reads)",
                      "reference_s": {"gfa_read ('Loaded graph' stamp)": rl, "gfa_read + read_gfa + reads ('Graph has' stamp, includes process start)": rt},
                      "phi_b200_s": {"phi_host_graph_load": round(g, 4), "phi_host_reads_load": round(r, 4), "sum": round(g + r, 4)},
                      "note": "reference: single thread; phi_b200 (round 2): a reader thread inflates while the caller's thread parses, W-line steps are resolved on up to 16 threads; zlib inflate of the GFA alone is ~0.07 s"}, indent=1))
