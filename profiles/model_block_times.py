#!/usr/bin/env python
"""SURVEY 8(f) row 3 measurement: model construction of the unmodified reference (oracle/_ref/PHI_ref) vs the integer-keyed
blocks of integration/phi_model.hpp (oracle/_ref/PHI_gpu_model, front end's result fed from a file written from the CPU oracle so
that no GPU is needed), from the reference's own log stamps.  Both against the recording Gurobi stand-in, -t1: once recording
(the dumps must be identical), once with PHI_STUB_QUIET=1 (the stand-in only counts the calls: what is left is the caller's work).
usage: model_block_times.py [backbone_bp haplotypes coverage]   (default: the README fixture shape through tests/golden)"""
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import phi_io  # noqa: E402
from phi_b200 import synth  # noqa: E402


def stamps(err):
    out = {}
    for line in err.splitlines():
        m = re.match(r"\[M::ILP_function::([\d.]+)\*[\d.]+\] (.*)", line)
        if m:
            out[m.group(2).split(":")[0].strip()] = float(m.group(1))
    return out


def main():
    if len(sys.argv) >= 4:
        sg = synth.make_graph(77, int(sys.argv[1]), int(sys.argv[2]))
        graph, reads, name = sg.graph, synth.make_reads(77, sg, float(sys.argv[3])), f"synthetic {sys.argv[1]} bp x {sys.argv[2]} haplotypes, {sys.argv[3]}x reads"
    else:
        from golden_cases import Case
        c = Case("mhc4")
        graph, reads, name = c.graph, c.reads, "README fixture (MHC, 5 walks, 111,805 vertices, 16,401 reads)"
    res = phi_io.oracle_index(graph, reads, 31, 25, 1.0)
    tmp = tempfile.mkdtemp()
    gfa, fa, rf = os.path.join(tmp, "g.gfa"), os.path.join(tmp, "r.fa"), os.path.join(tmp, "res.bin")
    synth.write_gfa(graph, gfa)
    synth.write_fasta(reads, fa)
    phi_io.write_result_file(rf, res, 2)
    rows = {}
    for q in ("1", "0"):
        sha = {}
        for exe, env in (("PHI_ref", {}), ("PHI_gpu_model", {"PHI_ADAPTER_RESULT_FILE": rf})):
            dump = os.path.join(tmp, f"{exe}_q{q}.dump")
            p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", exe), "-g", gfa, "-r", fa, "-o", dump + ".fa", "-t", "1", "-q", q],
                               env=dict(os.environ, PHI_STUB_DUMP=dump, **env), capture_output=True, text=True)
            assert p.returncode == 0, p.stderr[-1000:]
            s = stamps(p.stderr)
            start = s.get("QP model started", s.get("ILP model started"))
            rows[f"{exe} -q{q}"] = {"kmer_block_s": round(s["Minimizer constraints added to the model"] - start, 3),
                                    "expanded_graph_s": round(s["Optimized expanded graph constructed"] - s["Minimizer constraints added to the model"], 3)}
            sha[exe] = hashlib.sha256(open(dump, "rb").read()).hexdigest()
        assert sha["PHI_ref"] == sha["PHI_gpu_model"], "model dumps differ"
        # the same without the stand-in's text output (PHI_STUB_QUIET: calls are only counted): the caller's own share
        for exe, env in (("PHI_ref", {}), ("PHI_gpu_model", {"PHI_ADAPTER_RESULT_FILE": rf})):
            p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", exe), "-g", gfa, "-r", fa, "-o", os.path.join(tmp, "q.fa"), "-t", "1", "-q", q],
                               env=dict(os.environ, PHI_STUB_DUMP=os.path.join(tmp, "quiet.dump"), PHI_STUB_QUIET="1", **env), capture_output=True, text=True)
            assert p.returncode == 0, p.stderr[-1000:]
            s = stamps(p.stderr)
            start = s.get("QP model started", s.get("ILP model started"))
            rows[f"{exe} -q{q}"]["kmer_block_quiet_s"] = round(s["Minimizer constraints added to the model"] - start, 3)
            rows[f"{exe} -q{q}"]["expanded_graph_quiet_s"] = round(s["Optimized expanded graph constructed"] - s["Minimizer constraints added to the model"], 3)
    print(json.dumps({"workload": name, "anchors": int(res.n_anchors), "spectrum": int(res.count_sp_r), "identical_model_dump": True, "seconds": rows}, indent=1))


if __name__ == "__main__":
    main()
