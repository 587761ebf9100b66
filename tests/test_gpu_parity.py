"""GPU (B200): the CUDA path, called through the C ABI, against the CPU oracle and the golden fixtures.
Bit-exact: integer / byte / index work only (the single float compare, threshold*num_walks, is replicated)."""
import numpy as np
import pytest

import phi_io
import phi_b200
from phi_b200 import synth, _abi
from golden_cases import Case, SMALL, MHC, check_against_golden, assert_same_result
from test_oracle_golden import KAT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    ix = phi_b200.PhiGpuIndex()
    yield ix
    ix.close()


def test_device_murmur_known_answers(gpu):
    for key, want in KAT.items():
        assert int(gpu.hash128_to_64(key, len(key))[0]) == want, key
    rng = np.random.default_rng(7)
    for ln in list(range(1, 50)) + [63, 64, 65, 100, 255]:   # packed path (<= 32), word path and byte path, clean and dirty keys
        keys = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (64, ln))]
        keys[32:, rng.integers(0, ln)] = ord("N")
        got = gpu.hash128_to_64(keys.tobytes(), ln)
        want = [phi_io.oracle_hash(bytes(r)) for r in keys]
        assert got.tolist() == want, ln


@pytest.mark.parametrize("name", SMALL + MHC)
def test_index_matches_oracle_and_reference(gpu, name):
    c = Case(name)
    got = gpu.run(c.graph, c.reads, c.k, c.w, c.T)
    check_against_golden(c, got)
    want = phi_io.oracle_index(c.graph, c.reads, c.k, c.w, c.T)
    assert_same_result(want, got)


@pytest.mark.parametrize("name", SMALL + ["mhc4"])
def test_walk_sketch_matches_reference_index_kmers(gpu, name):
    c = Case(name)
    res, hashes = gpu.sketch_walks(c.graph, c.k, c.w)
    assert res.minimizers_per_walk.tolist() == c.meta["minimizers_per_walk"]
    assert phi_io.sha(hashes) == c.digests["wm_hash"]
    assert phi_io.sha(res.anchor_off) == c.digests["wm_voff"]
    assert phi_io.sha(res.anchor_vtx) == c.digests["wm_vtx"]


@pytest.mark.parametrize("seed,k,w,T,kw", [
    (101, 31, 25, 1.0, {}),
    (102, 31, 25, 0.6, dict(lower_frac=0.05, n_frac=0.01, r_lower_frac=0.05, r_n_frac=0.01)),
    (103, 17, 7, 1.0, dict(chop=5, var_spacing=15)),
    (104, 32, 31, 0.8, dict(chop=1000)),
    (105, 5, 3, 1.0, dict(var_spacing=10, chop=3)),
    (106, 28, 256, 1.0, {}),
    (107, 1, 1, 1.0, dict(snv_only=True)),
    # k > 32: no 2-bit packing, every k-mer is compared and hashed from its spelling (the reference takes any k, ILP_index.cpp:390-394)
    (108, 33, 25, 1.0, {}),
    (109, 64, 9, 0.7, dict(lower_frac=0.02, n_frac=0.005, r_lower_frac=0.02, r_n_frac=0.005)),
    (110, 101, 31, 1.0, dict(chop=12)),
])
def test_random_graphs_match_oracle(gpu, seed, k, w, T, kw):
    rk = {a[2:]: kw.pop(a) for a in list(kw) if a.startswith("r_")}
    sg = synth.make_graph(seed, 150000, 7, **kw)
    rd = synth.make_reads(seed, sg, 2.0, **rk)
    got = gpu.run(sg.graph, rd, k, w, T)
    want = phi_io.oracle_index(sg.graph, rd, k, w, T)
    assert_same_result(want, got)


def test_edge_cases(gpu):
    sg = synth.make_graph(5, 3000, 3)
    g = sg.graph
    empty_reads = phi_b200.Reads(np.zeros(1, dtype=np.uint64), np.zeros(0, dtype=np.uint8))
    for reads in (empty_reads,
                  phi_b200.Reads.from_strings(["ACGT", "", "ACGTACGTAC" * 3, ""]),          # all shorter than w+k-1
                  phi_b200.Reads.from_strings(["", "A" * 200, "", "acgtn" * 40, "T" * 55, "G" * 54])):
        got = gpu.run(g, reads, 31, 25, 1.0)
        want = phi_io.oracle_index(g, reads, 31, 25, 1.0)
        assert_same_result(want, got)
    # no walks / walks shorter than one window
    g0 = phi_b200.Graph(g.seg_off, g.seg_bases, np.zeros(1, dtype=np.uint64), np.zeros(0, dtype=np.uint32), g.top_order_map)
    rd = synth.make_reads(5, sg, 2.0)
    assert_same_result(phi_io.oracle_index(g0, rd), gpu.run(g0, rd))
    g1 = phi_b200.Graph(g.seg_off, g.seg_bases, np.array([0, 1, 1, 3], dtype=np.uint64), g.walk_vtx[:3], g.top_order_map)
    assert_same_result(phi_io.oracle_index(g1, rd), gpu.run(g1, rd))


def test_zero_length_segments_and_unsorted_top_order(gpu):
    sg = synth.make_graph(9, 40000, 4)
    g = sg.graph
    rd = synth.make_reads(9, sg, 3.0)
    # insert empty segments: every 7th vertex gets an empty twin that the walks also visit
    so = g.seg_off.astype(np.int64)
    n = g.n_vtx
    twin = np.arange(0, n, 7)
    seg_off2 = np.concatenate([so, np.full(len(twin), so[-1])]).astype(np.uint64)
    walk = g.walk_vtx.astype(np.int64)
    is_t = (walk % 7) == 0
    reps = np.where(is_t, 2, 1)
    walk2 = np.repeat(walk, reps)
    first = np.r_[True, walk2[1:] != walk2[:-1]] | ~np.repeat(is_t, reps)
    walk2 = np.where(first, walk2, n + walk2 // 7)
    cs = np.concatenate([[0], np.cumsum(reps)])
    walk_off2 = cs[g.walk_off.astype(np.int64)]
    top2 = np.concatenate([g.top_order_map, np.zeros(len(twin), dtype=np.int32)])
    g2 = phi_b200.Graph(seg_off2, g.seg_bases, walk_off2, walk2.astype(np.uint32), top2)
    assert_same_result(phi_io.oracle_index(g2, rd), gpu.run(g2, rd))
    # a permuted (still valid per-walk strictly monotone is NOT guaranteed) top order: reversed ids -> slow anchor path
    g3 = phi_b200.Graph(g.seg_off, g.seg_bases, g.walk_off, g.walk_vtx, (n - 1 - np.arange(n)).astype(np.int32))
    assert_same_result(phi_io.oracle_index(g3, rd), gpu.run(g3, rd))


def test_unsupported_parameters_fail_loudly(gpu):
    sg = synth.make_graph(5, 3000, 2)
    rd = synth.make_reads(5, sg, 1.0)
    with pytest.raises(phi_b200.PhiGpuError) as e:
        gpu.run(sg.graph, rd, 256, 25, 1.0)                 # vertex lists carry a one-byte length: k <= 255
    assert e.value.code == 2
    with pytest.raises(phi_b200.PhiGpuError):
        gpu.run(sg.graph, rd, 31, 300, 1.0)


def test_resident_run_equals_host_run_and_is_repeatable(gpu):
    c = Case("synth_small")
    a = gpu.run(c.graph, c.reads, c.k, c.w, c.T)
    gpu.upload(c.graph, c.reads)
    b = gpu.run_resident(c.k, c.w, c.T)
    b2 = gpu.run_resident(c.k, c.w, c.T)
    assert_same_result(a, b)
    assert_same_result(a, b2)
    t = gpu.times()
    assert t["kernel_launches"] > 10 and t["walk_kernel_ms"] > 0


def test_full_size_properties(gpu):
    """BASELINE configs[1] shape at reduced haplotype count (oracle too slow at full size): size-independent properties."""
    sg = synth.make_graph(0x50484931 + 1, 1_000_000, 12)
    rd = synth.make_reads(0x50484931 + 1, sg, 5.0)
    got = gpu.run(sg.graph, rd)
    assert np.all(np.diff(got.spectrum.astype(np.float64)) >= 0) and len(np.unique(got.spectrum)) == got.count_sp_r
    key = got.anchor_rank.astype(np.int64) * 1000 + got.anchor_walk
    assert np.all(np.diff(key) >= 0)                                   # sorted by (rank, walk)
    assert np.bincount(got.anchor_walk, minlength=12).tolist() == got.anchors_per_walk.tolist()
    # walk shuffling permutes per-walk results and nothing else
    perm = np.random.default_rng(1).permutation(12)
    wo = sg.graph.walk_off.astype(np.int64)
    walks = [sg.graph.walk_vtx[wo[h]:wo[h + 1]] for h in perm]
    g2 = phi_b200.Graph(sg.graph.seg_off, sg.graph.seg_bases, np.concatenate([[0], np.cumsum([len(x) for x in walks])]),
                        np.concatenate(walks), sg.graph.top_order_map)
    got2 = gpu.run(g2, rd)
    assert got2.minimizers_per_walk.tolist() == got.minimizers_per_walk[perm].tolist()
    assert got2.anchors_per_walk.tolist() == got.anchors_per_walk[perm].tolist()
    assert got2.n_filtered == got.n_filtered and np.array_equal(got2.spectrum, got.spectrum)
    # read order does not matter; duplicated reads change nothing (set semantics, ILP_index.cpp:631-635)
    rd2 = phi_b200.Reads(np.concatenate([rd.read_off, rd.read_off[1:] + rd.read_off[-1]]), np.concatenate([rd.read_bases, rd.read_bases]))
    got3 = gpu.run(sg.graph, rd2)
    assert np.array_equal(got3.spectrum, got.spectrum) and np.array_equal(got3.anchor_vtx, got.anchor_vtx)


@pytest.mark.parametrize("shift,share", [(4, True), (7, True), (11, False), (16, True), (24, True)])
def test_walk_sharing_does_not_change_results(gpu, shift, share):
    """Chunk boundaries and sharing only change the amount of work: tiny chunks (shorter than the w / k-1 context),
    no sharing at all, chunks longer than a tile, one chunk per walk — all bit-identical to the oracle."""
    sg = synth.make_graph(211, 120000, 9, lower_frac=0.01, n_frac=0.002)
    g = sg.graph
    # walk 3 is repeated verbatim (fully shared), walk 5 is cut short (shares a prefix only, ends inside a chunk)
    wo = g.walk_off.astype(np.int64)
    walks = [g.walk_vtx[wo[h]:wo[h + 1]] for h in range(g.n_walks)]
    walks.append(walks[3].copy())
    walks.append(walks[5][:len(walks[5]) // 3])
    walks.append(walks[5][len(walks[5]) // 2:])
    g2 = phi_b200.Graph(g.seg_off, g.seg_bases, np.concatenate([[0], np.cumsum([len(x) for x in walks])]),
                        np.concatenate(walks), g.top_order_map)
    rd = synth.make_reads(211, sg, 3.0, lower_frac=0.01, n_frac=0.002)
    want = phi_io.oracle_index(g2, rd, 31, 25, 0.7)
    try:
        gpu.set_walk_sharing(shift, share)
        got = gpu.run(g2, rd, 31, 25, 0.7)
        st = gpu.sharing()
        res, hashes = gpu.sketch_walks(g2, 21, 11)
    finally:
        gpu.set_walk_sharing()
    assert_same_result(want, got)
    assert st["unique_windows"] <= got.path_kmer_positions and st["unique_hits"] <= got.path_hits
    if share and shift <= 11:
        assert st["unique_windows"] < 0.7 * got.path_kmer_positions      # 9 walks from 8 founders + a verbatim copy: real sharing
    if not share:
        assert st["unique_windows"] == got.path_kmer_positions and st["unique_hits"] == got.path_hits
    ref, ref_hashes = gpu.sketch_walks(g2, 21, 11)
    assert np.array_equal(hashes, ref_hashes) and np.array_equal(res.anchor_vtx, ref.anchor_vtx)
    assert np.array_equal(res.anchor_off, ref.anchor_off) and np.array_equal(res.anchor_walk, ref.anchor_walk)
    assert res.minimizers_per_walk.tolist() == ref.minimizers_per_walk.tolist()


@pytest.mark.parametrize("name", ["toy_k3_w2", "synth_small", "synth_dirty", "synth_repeats"])
def test_debug_shared_kmer_statistic(gpu, name):
    """params.debug (-d1): distinct walk-minimizer hashes by the number of walks they occur in (ILP_index.cpp:565-606)."""
    c = Case(name)
    got = gpu.run(c.graph, c.reads, c.k, c.w, c.T, debug=1)
    want = phi_io.oracle_index(c.graph, c.reads, c.k, c.w, c.T, debug=1)
    assert_same_result(want, got)
    assert got.shared_kmer_hist is not None and got.n_walk_kmers == want.n_walk_kmers > 0
    assert got.shared_kmer_hist.tolist() == want.shared_kmer_hist.tolist()
    assert int(got.shared_kmer_hist.sum()) == got.n_walk_kmers and got.shared_kmer_hist[0] == 0
    plain = gpu.run(c.graph, c.reads, c.k, c.w, c.T)
    assert plain.shared_kmer_hist is None
    assert_same_result(plain, got)


@pytest.mark.parametrize("name", ["synth_small", "synth_dirty"])
def test_from_files_through_the_host_ingest(gpu, tmp_path, name):
    """GFA + FASTA on disk -> phi_host_graph_load / phi_host_reads_load -> GPU front end == oracle on the golden views."""
    c = Case(name)
    gfa, fa = str(tmp_path / "g.gfa"), str(tmp_path / "r.fa")
    synth.write_gfa(c.graph, gfa)
    synth.write_fasta(c.reads, fa)
    g = phi_b200.load_gfa(gfa)
    rd, _ = phi_b200.load_reads(fa)
    got = gpu.run(g, rd, c.k, c.w, c.T)
    want = phi_io.oracle_index(c.graph, c.reads, c.k, c.w, c.T)
    assert_same_result(want, got)


# ---- the other BASELINE.json configs as parity cases (shapes of configs[2..4] at a size the oracle finishes in seconds)
def test_baseline_config2_long_reads(gpu):
    """configs[2]: the MHC-shaped graph with 15 kb HiFi-like reads (log-normal lengths, 1 % errors) at 10x."""
    sg = synth.make_graph(0x50484931 + 2, 400_000, 12)
    rd = synth.make_reads(0x50484931 + 2, sg, 10.0, read_len=15000, len_sigma=0.2, sub_err=0.01)
    assert rd.n_reads > 100 and int(np.diff(rd.read_off.astype(np.int64)).max()) > 20000
    assert_same_result(phi_io.oracle_index(sg.graph, rd), gpu.run(sg.graph, rd))


def test_baseline_config3_vcf2gfa_shape_many_haplotypes(gpu):
    """configs[3]: vcf2gfa-shaped graph (SNV / small indel bubbles only, -m 30 chopping), 200 haplotypes, 150 bp reads at 10x."""
    sg = synth.make_graph(0x50484931 + 3, 150_000, 200, sv_frac=0.0, max_indel=20, founders=24, block_sites=150)
    rd = synth.make_reads(0x50484931 + 3, sg, 10.0)
    want = phi_io.oracle_index(sg.graph, rd, 31, 25, 1.0)
    got = gpu.run(sg.graph, rd, 31, 25, 1.0)
    assert_same_result(want, got)
    st = gpu.sharing()
    assert st["unique_windows"] < 0.5 * got.path_kmer_positions          # 200 walks from 24 founders per block: most chunks are shared
    # a fractional threshold on many walks exercises the float compare (count >= T * num_walks, ILP_index.cpp:698)
    assert_same_result(phi_io.oracle_index(sg.graph, rd, 31, 25, 0.335), gpu.run(sg.graph, rd, 31, 25, 0.335))


def test_baseline_config4_shape_500_haplotypes(gpu):
    """configs[4] shape: 500 haplotypes, 30x short reads (backbone scaled down so that the oracle finishes in seconds)."""
    sg = synth.make_graph(0x50484931 + 4, 40_000, 500, founders=32, block_sites=100)
    rd = synth.make_reads(0x50484931 + 4, sg, 30.0)
    assert_same_result(phi_io.oracle_index(sg.graph, rd), gpu.run(sg.graph, rd))


@pytest.mark.parametrize("n_haps", [2500, 4500])
def test_many_walks_group_members(gpu, n_haps):
    """More walks than the shared-memory paths of the grouped result hold (member merge by counting sort up to 2048 walks,
    block-local anchors-per-walk histogram up to 4096): the multi-pass merge and the global histogram give the same result."""
    sg = synth.make_graph(900 + n_haps, 3_000, n_haps, founders=12, block_sites=20)
    rd = synth.make_reads(900 + n_haps, sg, 20.0)
    for T in (1.0, 0.4):
        want = phi_io.oracle_index(sg.graph, rd, 31, 25, T)
        got = gpu.run(sg.graph, rd, 31, 25, T)
        assert_same_result(want, got)
    assert got.n_groups < got.n_anchors                                   # the lists are shared between walks
    # the members of every group ascend (checked on the raw arrays by result_to_py); most groups have several parts here
    assert int(np.diff(got.group_member_off.astype(np.int64)).max()) > 32


def test_more_than_65536_walks_use_32_bit_member_ids(gpu):
    """member_walk16 holds walk ids up to 65535; beyond that the result switches to member_walk32 (and the member merge runs one
    counting pass per 2048 walk ids)."""
    sg = synth.make_graph(977, 600, 70000, founders=6, block_sites=8, sv_frac=0.0)
    rd = synth.make_reads(977, sg, 30.0)
    want = phi_io.oracle_index(sg.graph, rd, 31, 25, 1.0)
    got = gpu.run(sg.graph, rd, 31, 25, 1.0)
    assert_same_result(want, got)
    assert got.member_walk_bytes == 4 and got.n_anchors > 100000 and int(got.member_walk.max()) > 65535
    sg2 = synth.make_graph(978, 20000, 5)
    small = gpu.run(sg2.graph, synth.make_reads(978, sg2, 5.0), 31, 25, 1.0)
    assert small.member_walk_bytes == 2 and small.n_anchors > 0


# ---- region partition of the walks on ONE GPU: every region is run by its own ctx (no communicator, so the group counts are
# local: with a threshold nothing can reach no rank is dropped and the merged parts must equal the oracle's unfiltered result)
@pytest.mark.parametrize("world,seed,kw", [(2, 31, {}), (3, 32, dict(lower_frac=0.01, n_frac=0.002)), (5, 33, dict(chop=8)), (4, 34, dict(founders=3, block_sites=25))])
def test_region_slices_merge_to_the_whole(gpu, world, seed, kw):
    from phi_b200 import multi
    sg = synth.make_graph(seed, 90_000, 9, **kw)
    rd = synth.make_reads(seed, sg, 4.0)
    g = sg.graph
    k, w, T = 31, 25, 1000.0
    want = phi_io.oracle_index(g, rd, k, w, T)
    bounds = multi.region_bounds(g, world)
    parts = []
    for r in range(world):
        gs = multi.slice_walks(g, k, w, bounds[r], bounds[r + 1])
        ix = phi_b200.PhiGpuIndex(0)
        ix.set_walk_region(bounds[r], bounds[r + 1])
        p = ix.run(gs, rd, k, w, T)
        ix.close()
        if r:
            p.spectrum = np.zeros(0, dtype=np.uint64)                      # as in a multi-GPU run: rank 0 alone carries the spectrum
            p.read_kmer_positions = p.read_minimizers_emitted = 0
        parts.append(p)
    got = multi.merge_results(parts)
    assert sum(p.path_kmer_positions for p in parts) == want.path_kmer_positions
    got.path_hits = want.path_hits
    assert_same_result(want, got)
    # the whole range in one region is the plain run
    ix = phi_b200.PhiGpuIndex(0)
    ix.set_walk_region(0, 2 ** 64 - 1)
    assert_same_result(want, ix.run(g, rd, k, w, T))
    ix.close()


def test_inconsistent_views_are_rejected(gpu):
    sg = synth.make_graph(5, 3000, 2)
    rd = synth.make_reads(5, sg, 1.0)
    g = sg.graph
    bad = phi_b200.Graph(g.seg_off, g.seg_bases, g.walk_off, g.walk_vtx.copy(), g.top_order_map)
    bad.walk_vtx[len(bad.walk_vtx) // 2] = g.n_vtx + 7                     # vertex id out of range: caught on the device by the step pass
    with pytest.raises(phi_b200.PhiGpuError) as e:
        gpu.run(bad, rd)
    assert e.value.code == _abi.PHI_ERR_ARG
    so = g.seg_off.copy(); so[3] = so[5] + 1                               # offsets that go backwards
    with pytest.raises(phi_b200.PhiGpuError) as e:
        gpu.run(phi_b200.Graph(so, g.seg_bases, g.walk_off, g.walk_vtx, g.top_order_map), rd)
    assert e.value.code == _abi.PHI_ERR_ARG
    ro = rd.read_off.copy(); ro[0] = 1
    with pytest.raises(phi_b200.PhiGpuError) as e:
        gpu.run(g, phi_b200.Reads(ro, rd.read_bases))
    assert e.value.code == _abi.PHI_ERR_ARG
    assert_same_result(phi_io.oracle_index(g, rd, 31, 25, 1.0), gpu.run(g, rd))   # the ctx is still usable
