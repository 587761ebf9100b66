"""CPU: the per-GPU views the reference-side adapter builds for a multi-GPU run (integration/phi_shards.hpp, compiled here into a
small harness over the library's host loaders) are exactly the shards the library's partition helpers define — the ones the
multi-GPU parity tests run on (phi_b200.multi.shard_inputs)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from phi_b200 import _abi, multi, synth
import phi_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "phi_b200")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("h") / "adapter_shards")
    cmd = ["g++", "-std=c++11", "-O1", "-fopenmp", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "integration"),
           os.path.join(ROOT, "tests", "adapter_shards_harness.cpp"), "-o", exe, "-L", LIBDIR, "-lphi_gpu_index", "-Wl,-rpath," + LIBDIR]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    return exe


def read_shards(path):
    b = open(path, "rb").read()
    at = [0]

    def u64(n=1):
        v = struct.unpack_from("<%dQ" % n, b, at[0]); at[0] += 8 * n
        return v if n > 1 else v[0]

    def arr(n, dt):
        a = np.frombuffer(b, dtype=dt, count=n, offset=at[0]); at[0] += a.nbytes
        return a

    world, by_region = u64(2)
    out = []
    for _ in range(world):
        region, lo, hi, base, n_vtx, n_walks, steps, n_reads, bases = u64(9)
        out.append(dict(region=region, lo=lo, hi=hi, base=base, n_vtx=n_vtx, walk_off=arr(n_walks + 1, np.uint64), walk_vtx=arr(steps, np.uint32),
                        read_off=arr(n_reads + 1, np.uint64), read_bases=arr(bases, np.uint8)))
    assert at[0] == len(b)
    return bool(by_region), out


def run_harness(exe, tmp_path, g, rd, world, k, w):
    gfa, fa, out = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa"), str(tmp_path / "s.bin")
    synth.write_gfa(g, gfa)
    synth.write_fasta(rd, fa)
    p = subprocess.run([exe, gfa, fa, str(world), str(k), str(w), out], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    g2 = phi_b200.load_gfa(gfa)                                    # vertex numbering of the loader (what the harness saw)
    rd2, _ = phi_b200.load_reads(fa)
    return g2, rd2, read_shards(out)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("k,w", [(31, 25), (5, 3)])
def test_adapter_shards_are_the_library_partition(harness, tmp_path, world, k, w):
    sg = synth.make_graph(11, 40000, 7, founders=3, block_sites=25)
    rd = synth.make_reads(11, sg, 4.0)
    g, rd, (by_region, shards) = run_harness(harness, tmp_path, sg.graph, rd, world, k, w)
    assert by_region == (world > 1) and len(shards) == world
    reads_seen = []
    for r, s in enumerate(shards):
        gs, rs, base, region = multi.shard_inputs(g, rd, r, world, k, w, "region")
        assert s["n_vtx"] == g.n_vtx
        assert np.array_equal(s["walk_off"], gs.walk_off.astype(np.uint64)) and np.array_equal(s["walk_vtx"], gs.walk_vtx)
        assert np.array_equal(s["read_off"], rs.read_off.astype(np.uint64)) and np.array_equal(s["read_bases"], rs.read_bases)
        assert s["base"] == base
        if world > 1:
            assert s["region"] == 1 and (s["lo"], s["hi"]) == (int(region[0]), int(region[1]))
        else:
            assert s["region"] == 0 and (s["lo"], s["hi"]) == (0, 2 ** 64 - 1)
        reads_seen.append(s["read_bases"])
    assert np.array_equal(np.concatenate(reads_seen), rd.read_bases)        # the read shards tile the read set


def test_adapter_falls_back_to_whole_walks(harness, tmp_path):
    """One more W line that runs AGAINST the links (a walk spelled backwards): no region cut is possible, the adapter hands out
    contiguous whole walks instead — the same ones phi_b200.multi.shard_inputs falls back to."""
    sg = synth.make_graph(12, 20000, 4, founders=2, block_sites=20)
    g0 = sg.graph
    rd = synth.make_reads(12, sg, 2.0)
    gfa, fa, out = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa"), str(tmp_path / "s.bin")
    synth.write_gfa(g0, gfa)
    wo = g0.walk_off.astype(np.int64)
    back = g0.walk_vtx[wo[1]:wo[2]][::-1].astype(np.int64) + 1
    with open(gfa, "a") as f:
        f.write("W\tbackwards\t0\tchr\t0\t0\t" + "".join(">s%d" % v for v in back) + "\n")
    synth.write_fasta(rd, fa)
    p = subprocess.run([harness, gfa, fa, "3", "31", "25", out], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    g2 = phi_b200.load_gfa(gfa)
    rd2, _ = phi_b200.load_reads(fa)
    assert g2.n_walks == g0.n_walks + 1
    by_region, shards = read_shards(out)
    assert not by_region
    walks_seen = 0
    for r, s in enumerate(shards):
        gs, rs, base, region = multi.shard_inputs(g2, rd2, r, 3, 31, 25, "region")
        assert region is None and s["region"] == 0 and s["base"] == base == walks_seen
        assert np.array_equal(s["walk_off"], gs.walk_off.astype(np.uint64)) and np.array_equal(s["walk_vtx"], gs.walk_vtx)
        walks_seen += len(s["walk_off"]) - 1
    assert walks_seen == g2.n_walks
