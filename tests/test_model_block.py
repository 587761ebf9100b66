"""CPU: SURVEY 8(f) row 3 — the k-mer constraint block of the model construction built straight from the grouped result and the
expanded graph of the default branch built with integer keys (integration/phi_model.hpp, replacing
/root/reference/src/ILP_index.cpp:782-880 and :1201-1406).  oracle/_ref/PHI_gpu_model is the reference CLI with the front end AND
those blocks replaced; here the front end's result comes from a file written from the CPU oracle's anchors
(test hook), so only the blocks are under test.  The recorded model must be the unmodified reference's, byte for byte: the dump of
oracle/_ref/PHI_ref on the same files, and for the synthetic cases also the SHA-256 tests/golden/ holds (the mhc4 fixture was
dumped from the original .gfa.gz, whose link order the re-written GFA does not keep)."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import phi_io
from phi_b200 import synth
from golden_cases import Case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "PHI_gpu_model")
REF = os.path.join(ROOT, "oracle", "_ref", "PHI_ref")


def test_grouping_round_trip():
    """group_anchors (test infrastructure) followed by the adapter's expansion gives the anchors back."""
    from phi_b200 import _abi
    c = Case("synth_repeats")
    want = phi_io.oracle_index(c.graph, c.reads, c.k, c.w, 2.0)
    rank_off, glen, gvtx, moff, mwalk = phi_io.group_anchors(want)
    grank = np.repeat(np.arange(want.count_sp_r, dtype=np.int32), np.diff(rank_off.astype(np.int64)))
    a_rank, a_walk, a_off, a_vtx = _abi.expand_groups(grank, glen, gvtx, moff, mwalk)
    assert np.array_equal(a_rank, want.anchor_rank) and np.array_equal(a_walk, want.anchor_walk)
    assert np.array_equal(a_off, want.anchor_off) and np.array_equal(a_vtx, want.anchor_vtx)
    assert len(glen) < want.n_anchors


@pytest.mark.parametrize("name,extra,member_bytes", [("toy_k3_w2", ["-k", "3", "-w", "2"], 2), ("synth_small", [], 2), ("synth_dirty", ["-T", "0.5"], 4),
                                                     ("synth_repeats", ["-T", "2.0"], 2), ("synth_unchopped", [], 2), ("mhc4", [], 2),
                                                     ("shape_200_haplotypes", ["-T", "0.335"], 2), ("shape_long_reads_15kb", [], 2)])
def test_model_block_from_the_grouped_result_dumps_the_reference_model(tmp_path, name, extra, member_bytes):
    if not (os.path.exists(EXE) and os.path.exists(REF)):
        pytest.skip("oracle/_ref/PHI_gpu_model / PHI_ref not built (it is built where /root/reference exists and travels with the snapshot)")
    c = Case(name)
    res = phi_io.oracle_index(c.graph, c.reads, c.k, c.w, c.T)
    gfa, fa, rf = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa"), str(tmp_path / "result.bin")
    synth.write_gfa(c.graph, gfa)
    synth.write_fasta(c.reads, fa)
    phi_io.write_result_file(rf, res, member_bytes)
    for q in ("1", "0"):
        dump = str(tmp_path / f"model_q{q}.dump")
        env = dict(os.environ, PHI_STUB_DUMP=dump, PHI_ADAPTER_RESULT_FILE=rf)
        p = subprocess.run([EXE, "-g", gfa, "-r", fa, "-o", dump + ".fa", "-t", "4", "-q", q] + extra, env=env, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-2000:]
        assert ("QP model started" if q == "1" else "ILP model started") in p.stderr
        got = hashlib.sha256(open(dump, "rb").read()).hexdigest()
        ref_dump = str(tmp_path / f"ref_q{q}.dump")
        p = subprocess.run([REF, "-g", gfa, "-r", fa, "-o", ref_dump + ".fa", "-t", "8", "-q", q] + extra, env=dict(os.environ, PHI_STUB_DUMP=ref_dump),
                           capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-2000:]
        assert got == hashlib.sha256(open(ref_dump, "rb").read()).hexdigest(), f"model dump differs from the reference's for -q{q}"
        if name != "mhc4":
            assert got == c.meta[f"model_q{q}_sha256"], f"model dump differs from the golden one for -q{q}"


@pytest.mark.parametrize("n_haps,T,flags", [(12, 1.0, []), (105, 0.3, []), (12, 1.0, ["-m", "0", "-R", "7"]), (12, 1.0, ["-N", "1"]),
                                            (30, 0.3, ["-N", "1", "-m", "0", "-R", "3"])])
def test_model_blocks_with_multi_digit_walk_ids(tmp_path, n_haps, T, flags):
    """The predecessor lists of the expanded graph come in the std::string order of names like "17_10" < "17_2" < "17_9": walk ids of
    two and three digits exercise the integer comparator that stands in for it.  Also: integer instead of mixed variables with an odd
recombination penalty (-m0 -R7: the coefficient is the integer c_1 / 2), and the naive expanded graph (-N1, ILP_index.cpp:942-1154,
replaced by add_naive_graph: same-walk and two-walk edge variables through integer-keyed tables)."""
    if not (os.path.exists(EXE) and os.path.exists(REF)):
        pytest.skip("oracle/_ref/PHI_gpu_model / PHI_ref not built")
    sg = synth.make_graph(4000 + n_haps, 4000, n_haps, founders=5, block_sites=12)
    rd = synth.make_reads(4000 + n_haps, sg, 8.0)
    res = phi_io.oracle_index(sg.graph, rd, 31, 25, T)
    assert res.n_anchors > 0
    gfa, fa, rf = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa"), str(tmp_path / "result.bin")
    synth.write_gfa(sg.graph, gfa)
    synth.write_fasta(rd, fa)
    phi_io.write_result_file(rf, res, 2)
    for q in ("1", "0"):
        dumps = []
        for exe, env in ((EXE, dict(PHI_ADAPTER_RESULT_FILE=rf)), (REF, {})):
            dump = str(tmp_path / f"{os.path.basename(exe)}_q{q}.dump")
            p = subprocess.run([exe, "-g", gfa, "-r", fa, "-o", dump + ".fa", "-t", "4", "-q", q, "-T", str(T)] + flags,
                               env=dict(os.environ, PHI_STUB_DUMP=dump, **env), capture_output=True, text=True)
            assert p.returncode == 0, p.stderr[-2000:]
            dumps.append(hashlib.sha256(open(dump, "rb").read()).hexdigest())
        assert dumps[0] == dumps[1], f"model dump differs from the reference's for -q{q}"


def test_integer_comparators_and_tables_against_std_string_and_std_map(tmp_path):
    """tests/cpp/model_order_selftest.cpp: dec_cmp / xnode_less (integration/phi_model_tables.hpp) order numbers and node names exactly as
    their std::strings compare; EdgeVarTable / PairIndex behave like std::map."""
    exe = str(tmp_path / "selftest")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "integration"),
                           os.path.join(ROOT, "tests", "cpp", "model_order_selftest.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "MODEL_ORDER_OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("bounds", [[0, 5, 12], [0, 1, 2, 7, 12], [0, 0, 12, 12]])
def test_model_blocks_from_several_parts(tmp_path, bounds):
    """Several GPUs return one result each (the groups of their own walks for all ranks); the adapter walks the parts one after another.
    Here the parts are cut out of the CPU oracle's anchors by walk range (also with empty parts) and fed through the file hook."""
    if not (os.path.exists(EXE) and os.path.exists(REF)):
        pytest.skip("oracle/_ref/PHI_gpu_model / PHI_ref not built")
    sg = synth.make_graph(4012, 4000, 12, founders=5, block_sites=12)
    rd = synth.make_reads(4012, sg, 8.0)
    res = phi_io.oracle_index(sg.graph, rd, 31, 25, 1.0)
    gfa, fa = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa")
    synth.write_gfa(sg.graph, gfa)
    synth.write_fasta(rd, fa)
    files = []
    for i, part in enumerate(phi_io.split_result_by_walks(res, bounds)):
        files.append(str(tmp_path / f"part{i}.bin"))
        phi_io.write_result_file(files[-1], part, 2 if i % 2 == 0 else 4)
    for q in ("1", "0"):
        outs = []
        for exe, env in ((EXE, dict(PHI_ADAPTER_RESULT_FILE=",".join(files))), (REF, {})):
            dump = str(tmp_path / f"{os.path.basename(exe)}_q{q}.dump")
            p = subprocess.run([exe, "-g", gfa, "-r", fa, "-o", dump + ".fa", "-t", "4", "-q", q],
                               env=dict(os.environ, PHI_STUB_DUMP=dump, **env), capture_output=True, text=True)
            assert p.returncode == 0, p.stderr[-2000:]
            counts = sorted(l for l in p.stderr.splitlines() if " : " in l or "Filtered/Retained" in l.split("]")[-1])
            counts = [l.split("] ")[-1] for l in counts]
            outs.append((hashlib.sha256(open(dump, "rb").read()).hexdigest(), counts))
        assert outs[0][0] == outs[1][0], f"model dump differs from the reference's for -q{q}"
        assert outs[0][1] == outs[1][1], "scraped per-walk counters differ"


@pytest.mark.parametrize("seed", range(8))
def test_group_and_expand_are_inverse_on_random_results(seed):
    """phi_b200._abi.expand_groups is what every GPU parity test looks through, phi_io.group_anchors what the CPU model tests feed the
    adapter with: on random anchors (repeated lists, a walk carrying a list twice, vertex ids whose decimal strings are prefixes of one
    another) grouping and expanding must give the anchors back, in (rank, walk, key order) order."""
    from phi_b200 import _abi
    rng = np.random.default_rng(seed)
    n_ranks, n_walks = int(rng.integers(1, 30)), int(rng.integers(1, 120))
    pool = [tuple(int(v) for v in rng.choice([1, 10, 100, 11, 2, 20, 9, 90, 99, 12345, 1234, 7], size=int(rng.integers(1, 5)))) for _ in range(12)]
    anchors = []                                                           # (rank, walk, list), reference order: key string, then insertion
    for r in range(n_ranks):
        per_walk = {}
        for _ in range(int(rng.integers(0, 40))):
            per_walk.setdefault(int(rng.integers(0, n_walks)), []).append(pool[int(rng.integers(0, len(pool)))])
        for h in sorted(per_walk):
            for lst in sorted(per_walk[h], key=lambda t: "".join(f"{v}_" for v in t).encode()):
                anchors.append((r, h, lst))
    lens = np.array([len(a[2]) for a in anchors], dtype=np.int64)
    res = _abi.IndexResultPy(
        count_sp_r=n_ranks, n_walks=n_walks, n_filtered=0, spectrum=np.arange(n_ranks, dtype=np.uint64),
        anchor_rank=np.array([a[0] for a in anchors], dtype=np.int32), anchor_walk=np.array([a[1] for a in anchors], dtype=np.int32),
        anchor_off=np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64),
        anchor_vtx=np.array([v for a in anchors for v in a[2]], dtype=np.int32),
        minimizers_per_walk=np.zeros(n_walks, dtype=np.uint64), anchors_per_walk=np.zeros(n_walks, dtype=np.uint64))
    rank_off, glen, gvtx, moff, mwalk = phi_io.group_anchors(res)
    grank = np.repeat(np.arange(n_ranks, dtype=np.int32), np.diff(rank_off.astype(np.int64)))
    a_rank, a_walk, a_off, a_vtx = _abi.expand_groups(grank, glen, gvtx, moff, mwalk)
    assert np.array_equal(a_rank, res.anchor_rank) and np.array_equal(a_walk, res.anchor_walk)
    assert np.array_equal(a_off, res.anchor_off) and np.array_equal(a_vtx, res.anchor_vtx)
