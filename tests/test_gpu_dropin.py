"""GPU (B200): the drop-in proof.  oracle/_ref/PHI_gpu is the reference CLI whose front end
(/root/reference/src/ILP_index.cpp:543-743) was replaced by libphi_gpu_index.so through integration/phi_adapter.hpp;
oracle/_ref/PHI_ref is the unmodified reference.  Both are linked against the recording Gurobi stand-in, so the
serialized ILP / IQP model they dump must be byte-identical, as must the counters PHI's own scripts scrape."""
import hashlib
import os
import re
import subprocess

import pytest

from phi_b200 import synth
from golden_cases import Case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "PHI_ref")
GPU = os.path.join(ROOT, "oracle", "_ref", "PHI_gpu")
GPU_MODEL = os.path.join(ROOT, "oracle", "_ref", "PHI_gpu_model")      # + the k-mer constraint block of integration/phi_model.hpp


def run(exe, gfa, fa, dump, extra):
    env = dict(os.environ, PHI_STUB_DUMP=dump)
    p = subprocess.run([exe, "-g", gfa, "-r", fa, "-o", dump + ".fa", "-t", "8"] + extra, env=env, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stderr


def scraped(err):
    keep = []
    for line in err.splitlines():
        line = re.sub(r"^\[M::(\w+)::[\d.]+\*[\d.]+\]", r"[M::\1]", line)
        if re.search(r"spectrum size|Filtered/Retained|Minimizers are in ILP| : \d+$|Number of|Haplotypes: \d+, fraction|Shared fraction", line):
            keep.append(line)
    head = [l for l in keep if "Number of Minimizers" not in l]
    return sorted(head)           # the reference prints the per-walk minimizer lines in thread order (ILP_index.cpp:563)


@pytest.mark.parametrize("name,extra", [("toy_k3_w2", ["-k", "3", "-w", "2"]), ("synth_small", []), ("synth_dirty", ["-T", "0.5"]),
                                        ("synth_small", ["-d", "1"]),
                                        ("synth_repeats", ["-T", "2.0"]), ("mhc4", [])])
def test_patched_reference_dumps_the_identical_model(tmp_path, name, extra):
    if not (os.path.exists(REF) and os.path.exists(GPU)):
        pytest.skip("oracle/_ref/PHI_ref / PHI_gpu not built (they are built where /root/reference exists and travel with the snapshot)")
    c = Case(name)
    gfa, fa = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa")
    synth.write_gfa(c.graph, gfa)
    synth.write_fasta(c.reads, fa)
    for q in ("1", "0"):
        d_ref, d_gpu = str(tmp_path / f"ref_q{q}.dump"), str(tmp_path / f"gpu_q{q}.dump")
        e_ref = run(REF, gfa, fa, d_ref, extra + ["-q", q])
        h_ref = hashlib.sha256(open(d_ref, "rb").read()).hexdigest()
        assert os.path.getsize(d_ref) > 0
        for exe in (GPU, GPU_MODEL):
            if not os.path.exists(exe):
                continue
            e_gpu = run(exe, gfa, fa, d_gpu, extra + ["-q", q])
            h_gpu = hashlib.sha256(open(d_gpu, "rb").read()).hexdigest()
            assert h_ref == h_gpu, f"model dump of {os.path.basename(exe)} differs for -q{q}"
            assert scraped(e_ref) == scraped(e_gpu)


@pytest.mark.parametrize("name,extra", [("synth_small", []), ("synth_dirty", ["-T", "0.5"]), ("synth_repeats", []), ("mhc4", [])])
def test_patched_reference_on_several_gpus_dumps_the_identical_model(tmp_path, name, extra):
    """PHI_GPU_DEVICES=0,1[,..]: the adapter runs one ctx per GPU (host thread each), walks sharded by region, reads by bases, the
    parts merged by phi_index_result_merge — the model dump must not change."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    if not (os.path.exists(REF) and os.path.exists(GPU)):
        pytest.skip("oracle/_ref/PHI_ref / PHI_gpu not built")
    c = Case(name)
    gfa, fa = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa")
    synth.write_gfa(c.graph, gfa)
    synth.write_fasta(c.reads, fa)
    d_ref, d_gpu = str(tmp_path / "ref.dump"), str(tmp_path / "gpu.dump")
    e_ref = run(REF, gfa, fa, d_ref, extra)
    os.environ["PHI_GPU_DEVICES"] = ",".join(str(i) for i in range(min(n, 4)))
    try:
        for exe in (GPU, GPU_MODEL):
            e_gpu = run(exe, gfa, fa, d_gpu, extra)
            assert hashlib.sha256(open(d_ref, "rb").read()).hexdigest() == hashlib.sha256(open(d_gpu, "rb").read()).hexdigest(), os.path.basename(exe)
            assert scraped(e_ref) == scraped(e_gpu)
    finally:
        del os.environ["PHI_GPU_DEVICES"]
