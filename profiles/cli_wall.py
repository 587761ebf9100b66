#!/usr/bin/env python
"""Ingest -> result wall of the stand-alone C command (examples/phi_index_cli.c) on configs[1]-sized FILES: writes the c2 graph as GFA
and the c2 reads as FASTQ, each as a bgzip (BGZF) container and as a single-member gzip, builds the command and runs it twice per
flavour with PHI_CLI_TIMES=1 PHI_HOST_TIMES=1.  profiles/r2_cli_wall_c2.txt is the output of these runs on a B200 box (the files were
written in the build container and travelled with the snapshot: writing them takes longer than everything measured here).
    python profiles/cli_wall.py prepare <dir>      # write the four files + the binary (CPU only)
    python profiles/cli_wall.py run <dir>          # on the GPU box"""
import gzip
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def prepare(d):
    import numpy as np
    import bench
    from phi_b200 import synth
    from test_host_io import bgzf_bytes
    os.makedirs(d, exist_ok=True)
    sys.argv = ["bench.py", "--config", "c2"]
    g, rd = bench.Workload(bench.parse_args()).shard(0, 1)[:2]
    gfa = os.path.join(d, "c2.gfa")
    synth.write_gfa(g, gfa)
    text = open(gfa, "rb").read()
    os.remove(gfa)
    ro, rb = rd.read_off.astype(np.int64), rd.read_bases.tobytes()
    fq = b"".join(b"@read%d\n" % i + rb[ro[i]:ro[i + 1]] + b"\n+\n" + b"I" * int(ro[i + 1] - ro[i]) + b"\n" for i in range(rd.n_reads))
    for name, data in (("gfa", text), ("fq", fq)):
        with open(os.path.join(d, "c2.bgzf.%s.gz" % name), "wb") as f:
            f.write(bgzf_bytes(data))
        with open(os.path.join(d, "c2.plain.%s.gz" % name), "wb") as f:
            f.write(gzip.compress(data, 1))
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "phi_index_cli.c"),
                           "-o", os.path.join(d, "phi_index_cli"), "-L", os.path.join(ROOT, "phi_b200"), "-lphi_gpu_index",
                           "-Wl,-rpath,/root/repo/phi_b200"])


def run(d):
    env = dict(os.environ, PHI_CLI_TIMES="1", PHI_HOST_TIMES="1")
    for kind in ("bgzf", "plain"):
        for rep in (1, 2):
            t0 = time.time()
            p = subprocess.run([os.path.join(d, "phi_index_cli"), "-g", os.path.join(d, "c2.%s.gfa.gz" % kind), "-r", os.path.join(d, "c2.%s.fq.gz" % kind)],
                               env=env, capture_output=True, text=True)
            for line in p.stderr.splitlines():
                if " : " not in line:
                    print("[%s %d] %s" % (kind, rep, line))
            print("[%s %d] process wall %.3f s" % (kind, rep, time.time() - t0))
    print("host threads:", os.cpu_count())


if __name__ == "__main__":
    {"prepare": prepare, "run": run}[sys.argv[1]](sys.argv[2])
