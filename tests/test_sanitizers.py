"""CPU: race and memory checks of the library's threaded host code.  The loaders (parallel BGZF inflate, read files parsed in
parallel chunks, streamed parse behind a reader thread), the partition helpers and phi_index_result_merge are linked from their
sources into tests/san_host_driver.cpp, once with ThreadSanitizer and once with AddressSanitizer + UBSan, and run with thread
counts and chunk sizes that put borders everywhere.  Any sanitizer report fails the test."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

from phi_b200 import synth
from golden_cases import Case
from test_host_io import bgzf_bytes, READS_EDGE, MALFORMED_READS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "phi_b200", "csrc", f) for f in ("host_io.cpp", "merge.cpp")]


@pytest.fixture(scope="module")
def drivers(tmp_path_factory):
    d = tmp_path_factory.mktemp("san")
    shard = str(d / "shard.cpp")
    shutil.copy(os.path.join(ROOT, "phi_b200", "csrc", "shard.cu"), shard)       # host-only code in a .cu file
    # shard.cu includes the public header relative to its own place
    text = open(shard).read().replace('#include "../../include/phi_gpu_index.h"', '#include "%s"' % os.path.join(ROOT, "include", "phi_gpu_index.h"))
    open(shard, "w").write(text)
    procs = {}
    for name, flags in (("thread", ["-fsanitize=thread"]), ("address", ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"])):
        exe = str(d / ("drv_" + name))
        procs[name] = (exe, subprocess.Popen(["g++", "-std=c++17", "-O1", "-g", "-fno-omit-frame-pointer"] + flags + SRC + [shard, os.path.join(ROOT, "tests", "san_host_driver.cpp"),
                                              "-o", exe, "-lz", "-lpthread"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    out = {}
    for name, (exe, p) in procs.items():
        log, _ = p.communicate()
        if p.returncode != 0 and ("cannot find" in log or "unrecognized" in log or "not supported" in log):
            pytest.skip("this toolchain has no -fsanitize=%s runtime" % name)
        assert p.returncode == 0, log[-3000:]
        out[name] = exe
    return out


def run_clean(exe, args, env=None):
    e = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66", ASAN_OPTIONS="detect_leaks=1 exitcode=67", UBSAN_OPTIONS="halt_on_error=1")
    e.update(env or {})
    p = subprocess.run([exe] + args, capture_output=True, text=True, env=e)
    report = [l for l in (p.stdout + p.stderr).splitlines() if "Sanitizer" in l or "runtime error" in l]
    assert p.returncode in (0,) and not report, (p.returncode, (p.stdout + p.stderr)[-3000:])
    return p.stdout


@pytest.mark.parametrize("san", ["thread", "address"])
def test_threaded_merge_is_clean_and_exact(drivers, san):
    out = run_clean(drivers[san], ["merge"])
    assert out.count("identical to the whole") == 3 and "MISMATCH" not in out


@pytest.mark.parametrize("san", ["thread", "address"])
def test_threaded_loaders_and_partition_helpers_are_clean(drivers, san, tmp_path):
    c = Case("synth_small")
    files = []
    gfa = str(tmp_path / "g.gfa")
    synth.write_gfa(c.graph, gfa)
    text = open(gfa, "rb").read()
    files.append(gfa)
    for name, data in (("g.bgzf.gfa.gz", bgzf_bytes(text, 3000)), ("g.plain.gfa.gz", gzip.compress(text)), ("g.cut.gfa.gz", gzip.compress(text)[:2000])):
        with open(str(tmp_path / name), "wb") as f:
            f.write(data)
        files.append(str(tmp_path / name))
    ro = c.reads.read_off.astype(np.int64)
    fq = b"".join(b"@r%d\n" % i + bytes(c.reads.read_bases[ro[i]:ro[i + 1]]) + b"\n+\n" + b"@" * int(ro[i + 1] - ro[i]) + b"\n" for i in range(c.reads.n_reads))
    corpus = {k: v.encode() for k, v in READS_EDGE.items()}
    corpus.update(MALFORMED_READS)
    corpus["reads.fq"] = fq
    for name, data in sorted(corpus.items()):
        for suffix, blob in (("", data), (".bgzf.gz", bgzf_bytes(data, 500)), (".gz", gzip.compress(data))):
            p = str(tmp_path / (name + suffix))
            with open(p, "wb") as f:
                f.write(blob)
            files.append(p)
    ref = run_clean(drivers[san], ["load"] + files, {"PHI_HOST_INFLATE_THREADS": "1", "PHI_SHARD_THREADS": "1"})
    for env in ({"PHI_HOST_PARSE_CHUNK": "5", "PHI_HOST_INFLATE_THREADS": "7", "PHI_SHARD_THREADS": "5", "PHI_SHARD_PIECE": "64"},
                {"PHI_HOST_PARSE_CHUNK": "301", "PHI_HOST_INFLATE_THREADS": "3", "PHI_SHARD_THREADS": "16"}):
        assert run_clean(drivers[san], ["load"] + files, env) == ref          # and the same counts whatever the threading
    assert "reads.fq reads rc=0 n=%d bases=%d" % (c.reads.n_reads, int(ro[-1])) in ref


def test_whole_buffer_inflate_against_zlib_under_asan(tmp_path):
    """phi_b200/csrc/fast_inflate.h: every zlib level and strategy on nine kinds of text decodes to the same bytes with the same
    consumed length; cut and bit-flipped streams are rejected or decoded inside the buffers (ASan + UBSan watch the bounds)."""
    exe = str(tmp_path / "inflate_fuzz")
    p = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                        os.path.join(ROOT, "tests", "inflate_fuzz.cpp"), "-o", exe, "-lz"], capture_output=True, text=True)
    if p.returncode != 0 and ("cannot find" in p.stderr or "not supported" in p.stderr):
        pytest.skip("no -fsanitize=address runtime")
    assert p.returncode == 0, p.stderr[-3000:]
    out = run_clean(exe, ["6"])
    assert "inflate_raw == zlib on" in out
