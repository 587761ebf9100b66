"""CPU: pins oracle/ (the C restatement) against what the UNMODIFIED reference produced (tests/golden)."""
import numpy as np
import pytest

import phi_io
from golden_cases import Case, SMALL, MHC, SHAPES, LONG_K, check_against_golden

# hash128_to_64 known answers produced by the reference's own MurmurHash3.cpp
KAT = {b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA": 0xc76d5c1cf7227aee, b"ACGTACGTACGTACGTACGTACGTACGTACG": 0x8f3c55213ed8e5fb,
       b"TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT": 0xba81eb486ec44359, b"AAAAAAAATTAGTACCTAACATTATCCTTTC": 0xd6760789f9d23068,
       b"ACGTNACGTNACGTNACGTNACGTNACGTNA": 0xb5ef1d09a0faeb6a, b"AAA": 0x8fcc8947e36bc820, b"ATC": 0xf22cdf8021364240,
       b"ACGTACGTACGTACGT": 0x0478f92bd663d4ce, b"ACGTACGTACGTACGTA": 0x6497e6737522a218}


def test_murmur_known_answers():
    for key, want in KAT.items():
        assert phi_io.oracle_hash(key) == want


@pytest.mark.parametrize("name", SMALL + SHAPES + LONG_K + MHC)
def test_oracle_matches_reference(name):
    c = Case(name)
    res = phi_io.oracle_index(c.graph, c.reads, c.k, c.w, c.T)
    check_against_golden(c, res)
    sk, hashes = phi_io.oracle_sketch_walks(c.graph, c.k, c.w)
    assert phi_io.sha(hashes) == c.digests["wm_hash"]
    assert phi_io.sha(sk.anchor_off) == c.digests["wm_voff"]
    assert phi_io.sha(sk.anchor_vtx) == c.digests["wm_vtx"]
    if "wm_hash" in c.z:
        assert np.array_equal(hashes, c.z["wm_hash"])
        assert np.array_equal(sk.anchor_vtx, c.z["wm_vtx"])


def test_lowercase_reads_give_the_same_model():
    a, b = Case("mhc4"), Case("mhc4_lower")
    assert a.meta["model_q1_sha256"] == b.meta["model_q1_sha256"]
    assert a.meta["model_q0_sha256"] == b.meta["model_q0_sha256"]


def test_readme_pins():
    m = Case("mhc4").meta
    assert m["count_sp_r"] == 138834
    assert m["minimizers_per_walk"] == [471226, 483005, 474157, 479033, 479135]
    assert m["anchors_per_walk"] == [23673, 11762, 13757, 10374, 11677]
    assert m["model_q1_counts"] == {"V": 836026, "C": 455633, "Q": 20717}
    assert m["model_q0_counts"] == {"V": 836026, "C": 519940, "Q": 0}
    assert Case("mhc4_N75").meta["count_sp_r"] == 150160


@pytest.mark.parametrize("name", ["toy_k3_w2", "synth_small", "synth_repeats"])
def test_oracle_debug_statistic_matches_reference_d1(tmp_path, name):
    """-d1: the 'fraction of unique shared kmers' lines (ILP_index.cpp:593-606) of the UNMODIFIED reference vs the oracle."""
    import os
    import subprocess
    import numpy as np
    from phi_b200 import synth
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "PHI_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/PHI_ref not built (needs /root/reference)")
    c = Case(name)
    gfa, fa = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa")
    synth.write_gfa(c.graph, gfa)
    synth.write_fasta(c.reads, fa)
    p = subprocess.run([ref, "-g", gfa, "-r", fa, "-o", str(tmp_path / "o.fa"), "-t", "4", "-d", "1", "-k", str(c.k), "-w", str(c.w)],
                       env=dict(os.environ, PHI_STUB_DUMP=str(tmp_path / "dump.txt")), capture_output=True, text=True)
    lines = [l for l in p.stderr.splitlines() if "Haplotypes:" in l]
    want = phi_io.oracle_index(c.graph, c.reads, c.k, c.w, c.T, debug=1)
    mine = ["[Haplotypes: %d, fraction of unique shared kmers: %.5f]"
            % (i, float(np.float32(int(want.shared_kmer_hist[i])) / np.float32(want.n_walk_kmers))) for i in range(1, c.graph.n_walks + 1)]
    assert lines == mine and len(lines) == c.graph.n_walks
