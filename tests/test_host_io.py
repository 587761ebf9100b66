"""CPU: the from-scratch host ingest (phi_host_graph_load / phi_host_reads_load) against the
UNMODIFIED reference's own parsers (gfa_read + ILP_index::read_gfa, kseq + read_ip_reads) run through
oracle/_ref/ref_probe, on the reference's fixtures and on hand-written edge cases; and against the golden graphs."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import phi_io
import phi_b200
from phi_b200 import synth
from golden_cases import Case, SMALL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBE = os.path.join(ROOT, "oracle", "_ref", "ref_probe")
REFTEST = "/root/reference/test"
needs_probe = pytest.mark.skipif(not os.path.exists(PROBE), reason="oracle/_ref/ref_probe not built (needs /root/reference)")


def probe(gfa, reads, out):
    cmd = [PROBE, "-g", gfa, "-o", out, "--graph-only"] + (["-r", reads] if reads else [])
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-1000:]
    return phi_io.read_phiarr(out)


def assert_valid_topological_order(g):
    """every walk step goes up in top_order_map, and the map is a permutation (acyclic inputs)"""
    top = g.top_order_map.astype(np.int64)
    assert sorted(top.tolist()) == list(range(g.n_vtx))
    wo = g.walk_off.astype(np.int64)
    for h in range(g.n_walks):
        t = top[g.walk_vtx[wo[h]:wo[h + 1]].astype(np.int64)]
        assert np.all(np.diff(t) > 0)


def assert_same_graph(mine, ref, ref_names):
    for f in ("seg_off", "seg_bases", "walk_off", "walk_vtx"):
        assert np.array_equal(getattr(mine, f), getattr(ref, f)), f
    assert mine.walk_names == ref_names


@pytest.mark.parametrize("name", SMALL + ["mhc4"])
def test_gfa_loader_reproduces_the_golden_graphs(tmp_path, name):
    """the golden graph views were dumped by the reference's parser; write them as GFA text and read them back"""
    c = Case(name)
    gfa = str(tmp_path / "g.gfa")
    synth.write_gfa(c.graph, gfa)
    g = phi_b200.load_gfa(gfa)
    assert_same_graph(g, c.graph, c.graph.walk_names or [f"hap{h}.{h}" for h in range(c.graph.n_walks)])
    assert_valid_topological_order(g)
    # same front-end result with either topological order (oracle: CPU)
    if name != "mhc4":
        a = phi_io.oracle_index(c.graph, c.reads, c.k, c.w, c.T)
        b = phi_io.oracle_index(g, c.reads, c.k, c.w, c.T)
        assert np.array_equal(a.anchor_vtx, b.anchor_vtx) and np.array_equal(a.anchor_off, b.anchor_off) and np.array_equal(a.anchor_walk, b.anchor_walk)


@needs_probe
@pytest.mark.skipif(not os.path.isdir(REFTEST), reason="reference fixtures not present")
@pytest.mark.parametrize("gfa,reads", [("test.gfa", "read.fa"), ("MHC_4.gfa.gz", "CHM13_reads.fq.gz")])
def test_loaders_match_the_reference_parsers_on_its_fixtures(tmp_path, gfa, reads):
    d = probe(os.path.join(REFTEST, gfa), os.path.join(REFTEST, reads), str(tmp_path / "p.phiarr"))
    ref = phi_io.graph_from_arrays(d)
    g = phi_b200.load_gfa(os.path.join(REFTEST, gfa))
    assert_same_graph(g, ref, ref.walk_names)
    assert_valid_topological_order(g)
    rd, names = phi_b200.load_reads(os.path.join(REFTEST, reads))
    assert np.array_equal(rd.read_off, d["read_off"]) and np.array_equal(rd.read_bases, d["read_bases"])
    assert len(names) == rd.n_reads and all(names)


GFA_EDGE = "\n".join([
    "H\tVN:Z:1.1",
    "W\tearly\t3\tchr\t0\t0\t>a>b",                 # in front of the S-lines: neither step names a segment yet, both are dropped
    "S\ta\tACGTACGTAC",
    "L\ta\t+\tzz\t+\t0M",                         # zz first appears on an L-line: it gets id 1 before its S-line
    "S\tb\tGGGTTT\tLN:i:6\txx:Z:tag",
    "S\tzz\tCCCCC",
    "P\tignored\ta+,b+\t*",
    "S\tc\ttttgggaaac",                            # lower case is kept verbatim
    "L\tzz\t+\tb\t+\t0M",
    "L\tb\t+\tc\t+\t0M",
    "L\ta\t+\tb\t+\t0M",
    "L\ta\t+\tb\t+\t0M",                           # duplicate arc
    "x",                                           # short line
    "W\ts1\t0\tchr\t0\t0\t>a>zz>b>c",
    "W\ts1\t1\tchr\t0\t0\t<c<b<a",                 # reversed walk: flipped to >a>b>c
    "W\ts2\t7\tchr\t0\t0\t>a>nosuch>b",            # unknown segment dropped
    "W\ts3\t0\tchr\t0\t0\tXa>b",                   # gfa_parse_W takes the first token whatever its first byte is: name "a", forward
    "W\ts4\t0\tchr\t0\t0\tb>c",                    # ... here the first token's name is empty: dropped, the walk is >c
    "W\ts5\t0\tchr\t0\t0\t>a>c",                   # a -> c is backed by no L-line
    "W\tshort\t0\tchr",                            # too few fields: ignored
    "S\tlate\tAC",                                 # defined after the W-lines
    ""]) + "\n"


@needs_probe
@pytest.mark.parametrize("crlf,gz", [(False, False), (True, False), (False, True)])
def test_gfa_edge_cases_match_the_reference_parser(tmp_path, crlf, gz):
    text = GFA_EDGE.replace("\n", "\r\n") if crlf else GFA_EDGE
    path = str(tmp_path / ("g.gfa.gz" if gz else "g.gfa"))
    with (gzip.open(path, "wb") if gz else open(path, "wb")) as f:
        f.write(text.encode())
    d = probe(path, None, str(tmp_path / "p.phiarr"))
    ref = phi_io.graph_from_arrays(d)
    g = phi_b200.load_gfa(path)
    assert_same_graph(g, ref, ref.walk_names)
    assert g.segment_names[:2] == ["a", "zz"] and g.walk_names == ["early.3", "s1.0", "s1.1", "s2.7", "s3.0", "s4.0", "s5.0"]
    assert g.n_unlinked_steps == 1                                  # only s5's a -> c
    wo = g.walk_off.astype(int)
    assert wo[1] == 0                                               # the early walk is empty
    assert g.walk_vtx[wo[2]:wo[3]].tolist() == [0, 2, 3]           # the reversed walk came out forward


GFA_NAMES = "\n".join([
    "H\tVN:Z:1.1",
    "S\ts7\tACGTAC",                               # first <prefix><number> name: "s" becomes the prefix of the direct-index route
    "S\ts07\tGGGA",                                # leading zero: an ordinary name, not s7
    "S\t7\tTTTT",                                  # no prefix: ordinary
    "S\tx7\tCCAA",                                 # other prefix: ordinary
    "S\ts123456789\tAAC",                          # nine digits: ordinary
    "S\ts16777216\tGT",                            # beyond the direct range: ordinary
    "S\ts0\tA",
    "S\ts\tCG",                                    # no digits
    "S\ts7a\tTG",                                  # digits not at the end
    "L\ts7\t+\ts07\t+\t0M", "L\ts07\t+\t7\t+\t0M", "L\t7\t+\tx7\t+\t0M", "L\tx7\t+\ts123456789\t+\t0M",
    "L\ts123456789\t+\ts16777216\t+\t0M", "L\ts16777216\t+\ts0\t+\t0M", "L\ts0\t+\ts\t+\t0M", "L\ts\t+\ts7a\t+\t0M",
    "L\ts7a\t+\ts9\t+\t0M",                     # s9 first appears on an L-line
    "S\ts9\tACG",
    "W\th\t0\tc\t0\t0\t>s7>s07>7>x7>s123456789>s16777216>s0>s>s7a>s9",
    "W\th\t1\tc\t0\t0\t>s7>s007>s9",           # s007 is unknown: dropped
    ""]) + "\n"


@needs_probe
def test_gfa_segment_name_routes_match_the_reference_parser(tmp_path):
    """Names of the form <prefix><number> are looked up through a direct index, everything else through the hash table: the mix
    of both must number the segments exactly as the reference does."""
    path = str(tmp_path / "names.gfa")
    open(path, "w").write(GFA_NAMES)
    d = probe(path, None, str(tmp_path / "p.phiarr"))
    ref = phi_io.graph_from_arrays(d)
    g = phi_b200.load_gfa(path)
    assert_same_graph(g, ref, ref.walk_names)
    assert g.segment_names == ["s7", "s07", "7", "x7", "s123456789", "s16777216", "s0", "s", "s7a", "s9"]
    wo = g.walk_off.astype(int)
    assert g.walk_vtx[wo[0]:wo[1]].tolist() == list(range(10)) and g.walk_vtx[wo[1]:wo[2]].tolist() == [0, 9]


def test_gfa_reverse_strand_walk_is_an_error(tmp_path):
    path = str(tmp_path / "g.gfa")
    with open(path, "w") as f:
        f.write("S\ta\tACGT\nS\tb\tGGGG\nW\ts\t0\tc\t0\t0\t>a>b\nW\ts\t1\tc\t0\t0\t>a<b\n")
    with pytest.raises(phi_b200.PhiGpuError) as e:                  # ILP_index.cpp:104-107 exits there
        phi_b200.load_gfa(path)
    assert e.value.code == 2 and "reverse strand" in str(e.value)
    with pytest.raises(phi_b200.PhiGpuError):
        phi_b200.load_gfa(str(tmp_path / "missing.gfa"))


READS_EDGE = {
    "multi.fa": ">r1 comment here\nACGT\nacgtn\n\nGG\n>r2\n>r3\tx\nTTTT\n",
    "crlf.fq": "@q1\r\nACGTAC\r\n+\r\nIIIIII\r\n@q2 c\r\nGGG\r\nTT\r\n+q2\r\nII\r\nIII\r\n",
    "mixed.fq": "junk before\n@a\nACGT\n+\n@@@@\n>b\nCCCC\nGG\n@c\nTT\n+\n!!\n",
    "trunc.fq": "@ok\nACGT\n+\nIIII\n@bad\nACGTACGT\n+\nIII\n@never\nAC\n+\nII\n",
    "empty.fa": "",
    "noseq.fa": ">only_header\n",
}


@needs_probe
@pytest.mark.parametrize("fname", sorted(READS_EDGE))
def test_read_loader_matches_kseq(tmp_path, fname):
    gfa = str(tmp_path / "g.gfa")
    with open(gfa, "w") as f:
        f.write("S\ta\tACGT\n")
    path = str(tmp_path / fname)
    with open(path, "wb") as f:
        f.write(READS_EDGE[fname].encode())
    d = probe(gfa, path, str(tmp_path / "p.phiarr"))
    rd, names = phi_b200.load_reads(path)
    assert rd.read_off.tolist() == d["read_off"].tolist()
    assert bytes(rd.read_bases) == bytes(d["read_bases"])


def test_written_fasta_round_trip(tmp_path):
    c = Case("synth_small")
    fa = str(tmp_path / "r.fa.gz")
    plain = str(tmp_path / "r.fa")
    synth.write_fasta(c.reads, plain)
    with gzip.open(fa, "wb") as f, open(plain, "rb") as src:
        f.write(src.read())
    for path in (plain, fa):
        rd, names = phi_b200.load_reads(path)
        assert np.array_equal(rd.read_off, c.reads.read_off) and np.array_equal(rd.read_bases, c.reads.read_bases)
        assert names[:2] == ["r0", "r1"]


@needs_probe
@pytest.mark.parametrize("seed", range(40))
def test_random_gfas_match_the_reference_parser(tmp_path, seed):
    """Random acyclic GFAs with every naming style mixed (numbers, prefixed numbers, leading zeros, words), segments first seen on L-lines
    (S-lines after their uses), reversed walks, unknown walk steps, CRLF: same numbering, sequences and walks as the reference."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(5, 60))
    styles = [lambda i: str(i + 1), lambda i: f"s{i + 1}", lambda i: f"s{i + 1:04d}", lambda i: f"node_{i}", lambda i: f"{i + 1}x",
              lambda i: f"utg{i * 7 + 3}", lambda i: f"s{(i + 1) * 1000003}"]
    mix = rng.random() < 0.5
    pick = int(rng.integers(0, len(styles)))
    names = []
    for i in range(n):
        nm = styles[int(rng.integers(0, len(styles))) if mix else pick](i)
        names.append(nm if nm not in names else nm + f"_{i}")
    order = rng.permutation(n)                                        # topological order of the DAG: order[a] before order[b] for a < b
    edges = set()
    walks = []
    for _ in range(int(rng.integers(1, 6))):
        steps = sorted(rng.choice(n, size=int(rng.integers(2, min(n, 12) + 1)), replace=False).tolist())
        w = [int(order[s]) for s in steps]
        edges.update(zip(w[:-1], w[1:]))
        walks.append(w)
    lines = ["H\tVN:Z:1.1"]
    recs = [("S", i) for i in range(n)] + [("L", e) for e in edges]   # (the reference crashes on segments that are never defined: all are)
    rng.shuffle(recs)
    for kind, x in recs:
        if kind == "S":
            seq = "".join(rng.choice(list("ACGTacgtN"), size=int(rng.integers(1, 9))))      # (no '*': the reference crashes on it)
            lines.append(f"S\t{names[x]}\t{seq}" + ("\tLN:i:5" if rng.random() < 0.2 else ""))
        else:
            lines.append(f"L\t{names[x[0]]}\t+\t{names[x[1]]}\t+\t0M")
    for h, w in enumerate(walks):
        if rng.random() < 0.3:                                        # written backwards: the flip rule turns it forward again
            body = "".join("<" + names[v] for v in reversed(w))
        else:
            body = "".join(">" + names[v] for v in w)
        if rng.random() < 0.3:
            body += ">no_such_segment"
        lines.append(f"W\tsample{h % 2}\t{h}\tchr\t0\t0\t{body}")
    text = ("\r\n" if rng.random() < 0.25 else "\n").join(lines) + "\n"
    path = str(tmp_path / "rand.gfa")
    open(path, "w", newline="").write(text)
    p = subprocess.run([PROBE, "-g", path, "-o", str(tmp_path / "p.phiarr"), "--graph-only"], capture_output=True, text=True)
    if p.returncode == 1 and "reverse strand" in p.stderr:            # a backward walk set the strand of its segments: both refuse
        with pytest.raises(phi_b200.PhiGpuError):
            phi_b200.load_gfa(path)
        return
    assert p.returncode == 0, p.stderr[-1000:]
    ref = phi_io.graph_from_arrays(phi_io.read_phiarr(str(tmp_path / "p.phiarr")))
    g = phi_b200.load_gfa(path)
    assert_same_graph(g, ref, ref.walk_names)
    assert_valid_topological_order(g)


def test_read_loader_streams_and_falls_back(tmp_path):
    """The read loader parses while a reader thread inflates, into a buffer sized by the gzip trailer.  A multi-member gzip file
    promises only its last member's size: the streamed pass overflows and the loader must fall back to inflate-then-parse, with the
    same result as on the plain file (and as on a single-member file, where the promise holds)."""
    rng = np.random.default_rng(5)
    recs = []
    for i in range(4000):
        seq = "".join(rng.choice(list("ACGTN"), size=int(rng.integers(1, 400))))
        recs.append(f"@r{i} comment {i}\n{seq}\n+\n{'I' * len(seq)}\n" if i % 3 else f">f{i}\n{seq[:len(seq) // 2]}\n{seq[len(seq) // 2:]}\n")
    text = "".join(recs).encode()
    plain, single, multi = str(tmp_path / "r.fx"), str(tmp_path / "single.gz"), str(tmp_path / "multi.gz")
    open(plain, "wb").write(text)
    with gzip.open(single, "wb") as f:
        f.write(text)
    cut = len(text) * 3 // 4
    cut = text.index(b"\n", cut) + 1
    with open(multi, "wb") as f:                                   # two members: the trailer of the file describes the second one only
        f.write(gzip.compress(text[:cut]))
        f.write(gzip.compress(text[cut:]))
    want, want_names = phi_b200.load_reads(plain)
    assert want.n_reads == 4000
    for path in (single, multi):
        got, names = phi_b200.load_reads(path)
        assert np.array_equal(got.read_off, want.read_off) and np.array_equal(got.read_bases, want.read_bases) and names == want_names


def test_gfa_loader_streams_and_falls_back(tmp_path):
    """The same for the GFA loader: a two-member gzip file overflows the streamed pass, the scan is repeated over the text inflated
    the plain way, and W-lines (cut during the scan, resolved afterwards in parallel) still only see the segments defined above them."""
    c = Case("synth_small")
    plain = str(tmp_path / "g.gfa")
    synth.write_gfa(c.graph, plain)
    text = open(plain, "rb").read()
    # a W-line in front of every S-line: all its steps are unknown at that point and must be dropped (gfa-io.cpp:399-405)
    early = b"W\tearly\t0\tchr\t0\t0\t>s1>s2>s3\n"
    text = text.replace(b"S\ts1\t", early + b"S\ts1\t", 1)
    open(plain, "wb").write(text)
    cut = text.index(b"\n", len(text) // 2) + 1
    multi_gz, single_gz = str(tmp_path / "multi.gfa.gz"), str(tmp_path / "single.gfa.gz")
    with open(multi_gz, "wb") as f:
        f.write(gzip.compress(text[:cut]))
        f.write(gzip.compress(text[cut:]))
    with gzip.open(single_gz, "wb") as f:
        f.write(text)
    want = phi_b200.load_gfa(plain)
    wo = want.walk_off.astype(int)
    assert want.walk_names[0] == "early.0" and wo[1] == 0            # known at its line: nothing
    for path in (single_gz, multi_gz):
        got = phi_b200.load_gfa(path)
        assert_same_graph(got, want, want.walk_names)


MALFORMED_GFA = {
    "empty": b"", "no_newline": b"S\t1\tACGT", "s_one_field": b"S\n", "s_two_fields": b"S\t1\n", "s_empty_seq": b"S\t1\t\n",
    "l_short": b"S\t1\tACGT\nL\t1\n", "w_short": b"S\t1\tACGT\nW\tx\n", "w_unknown_segments": b"S\t1\tACGT\nW\ts\t0\tc\t0\t4\t>7>8\n",
    "w_empty": b"S\t1\tACGT\nW\ts\t0\tc\t0\t4\t\n", "w_star": b"S\t1\tACGT\nW\ts\t0\tc\t0\t4\t*\n", "w_bad_hap": b"S\t1\tACGT\nW\ts\tzz\tc\t0\t4\t>1\n",
    "cycle": b"S\t1\tACGT\nS\t2\tAC\nL\t1\t+\t2\t+\t0M\nL\t2\t+\t1\t+\t0M\nW\ts\t0\tc\t0\t4\t>1>2\n", "duplicate_s": b"S\t1\tACGT\nS\t1\tAC\n",
    "only_tabs": b"\t\t\t\n\t\nS\t\t\n", "long_line": b"S\t1\t" + b"A" * 3000000 + b"\n",
    "binary": bytes(np.random.default_rng(1).integers(0, 256, 5000, dtype=np.uint8)),
}
MALFORMED_READS = {
    "no_newline.fa": b">r\nACGT", "no_name.fa": b"ACGT\nACGT\n", "fq_cut_in_quality.fq": b"@r\nACGT\n+\nII", "fq_no_plus.fq": b"@r\nACGT\nIIII\n",
    "fq_short_quality.fq": b"@r\nACGT\n+\nI\n@r2\nAC\n+\nII\n", "blank_lines.fa": b">r\nAC\nGT\n\n>s\n\n>t\nA\n", "headers_only.fa": b">\n>\n>\n",
    "binary.fa": bytes(np.random.default_rng(2).integers(0, 256, 5000, dtype=np.uint8)),
}


@needs_probe
@pytest.mark.parametrize("name", sorted(MALFORMED_GFA))
def test_malformed_gfa_is_read_like_the_reference_reads_it(tmp_path, name):
    """Broken or degenerate GFA text: whatever the reference's parser makes of it (gfa-io.cpp skips what it cannot use), this one
    makes the same of it — and never crashes."""
    path = str(tmp_path / "m.gfa")
    with open(path, "wb") as f:
        f.write(MALFORMED_GFA[name])
    d = probe(path, None, str(tmp_path / "p.phiarr"))
    g = phi_b200.load_gfa(path)
    assert (g.n_vtx, g.n_walks) == (len(d["seg_off"]) - 1, len(d["walk_off"]) - 1)
    assert g.seg_off.tolist() == d["seg_off"].tolist() and bytes(g.seg_bases) == bytes(d["seg_bases"])
    assert g.walk_off.tolist() == d["walk_off"].tolist() and g.walk_vtx.tolist() == d["walk_vtx"].tolist()


@needs_probe
@pytest.mark.parametrize("fname", sorted(MALFORMED_READS))
def test_malformed_reads_are_read_like_kseq_reads_them(tmp_path, fname):
    gfa = str(tmp_path / "g.gfa")
    with open(gfa, "w") as f:
        f.write("S\ta\tACGT\n")
    path = str(tmp_path / fname)
    with open(path, "wb") as f:
        f.write(MALFORMED_READS[fname])
    d = probe(gfa, path, str(tmp_path / "p.phiarr"))
    rd, names = phi_b200.load_reads(path)
    assert rd.read_off.tolist() == d["read_off"].tolist() and bytes(rd.read_bases) == bytes(d["read_bases"])


@pytest.mark.parametrize("cut", [150, 1000])
def test_truncated_gzip_yields_what_could_be_inflated(tmp_path, cut):
    """A gzip file cut short: gzread (the reference's reader) hands out what inflates and then stops; so does this loader."""
    import gzip
    body = b"".join(b"@r%d\nACGTACGTAC\n+\nIIIIIIIIII\n" % i for i in range(5000))
    path = str(tmp_path / "t.fq.gz")
    with open(path, "wb") as f:
        f.write(gzip.compress(body)[:cut])
    rd, names = phi_b200.load_reads(path)
    assert 0 < rd.n_reads < 5000 and np.all(np.diff(rd.read_off.astype(np.int64))[:-1] == 10)
    if os.path.exists(PROBE):
        gfa = str(tmp_path / "g.gfa")
        with open(gfa, "w") as f:
            f.write("S\ta\tACGT\n")
        d = probe(gfa, path, str(tmp_path / "p.phiarr"))
        assert rd.read_off.tolist() == d["read_off"].tolist() and bytes(rd.read_bases) == bytes(d["read_bases"])


def bgzf_bytes(data, block=65280, level=6, eof=True):
    """bgzip's container: independent gzip members of <= 64 KB of text, each header carrying the member's size in a 'BC' extra subfield."""
    import struct
    import zlib
    out = bytearray()

    def member(chunk):
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        d = c.compress(chunk) + c.flush()
        out.extend(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(d) + 25) + d
                   + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
    for i in range(0, len(data), block):
        member(data[i:i + block])
    if eof:
        member(b"")
    return bytes(out)


@pytest.mark.parametrize("threads,block,eof", [(1, 65280, True), (5, 4096, True), (16, 700, False)])
def test_bgzf_files_are_inflated_in_parallel_to_the_same_text(tmp_path, monkeypatch, threads, block, eof):
    """bgzip-compressed GFA and FASTQ (the reference's own MHC_4.gfa.gz is one): the members are inflated by several threads while the
    parser follows the finished prefix — the outcome is the plain file's."""
    monkeypatch.setenv("PHI_HOST_INFLATE_THREADS", str(threads))
    c = Case("synth_small")
    gfa, fq = str(tmp_path / "g.gfa"), str(tmp_path / "r.fq")
    synth.write_gfa(c.graph, gfa)
    ro = c.reads.read_off.astype(np.int64)
    with open(fq, "wb") as f:
        for i in range(c.reads.n_reads):
            s = bytes(c.reads.read_bases[ro[i]:ro[i + 1]])
            f.write(b"@r%d some comment\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")
    for path in (gfa, fq):
        with open(path + ".gz", "wb") as f:
            f.write(bgzf_bytes(open(path, "rb").read(), block, 6, eof))
        assert gzip.decompress(open(path + ".gz", "rb").read()) == open(path, "rb").read()
    a, b = phi_b200.load_gfa(gfa), phi_b200.load_gfa(gfa + ".gz")
    for fld in ("seg_off", "seg_bases", "walk_off", "walk_vtx", "top_order_map"):
        assert np.array_equal(getattr(a, fld), getattr(b, fld)), fld
    assert a.walk_names == b.walk_names and a.segment_names == b.segment_names
    (ra, na), (rb, nb) = phi_b200.load_reads(fq), phi_b200.load_reads(fq + ".gz")
    assert np.array_equal(ra.read_off, rb.read_off) and np.array_equal(ra.read_bases, rb.read_bases) and na == nb
    assert np.array_equal(ra.read_bases, c.reads.read_bases)


def test_bgzf_look_alikes_take_the_serial_path(tmp_path):
    """Whatever is not a well-formed chain of BGZF members from the first to the last byte is read the way gzread reads it: a
    BGZF chain followed by an ordinary gzip member (concatenation = concatenated text), a member whose checksum is wrong (error,
    as before), a chain cut inside a member (what inflates is used)."""
    body = b"".join(b">r%d\n%s\n" % (i, b"ACGT" * 30) for i in range(3000))
    good = bgzf_bytes(body, 5000)
    mixed = str(tmp_path / "mixed.fa.gz")
    with open(mixed, "wb") as f:
        f.write(bgzf_bytes(body, 5000, eof=False) + gzip.compress(b">last\nGGGG\n"))
    rd, names = phi_b200.load_reads(mixed)
    assert rd.n_reads == 3001 and names[-1] == "last" and bytes(rd.read_bases[-4:]) == b"GGGG"
    bad = bytearray(good)
    bad[len(good) // 2] ^= 0x55                                              # somewhere inside a member's deflate data
    broken = str(tmp_path / "broken.fa.gz")
    with open(broken, "wb") as f:
        f.write(bytes(bad))
    with pytest.raises(phi_b200.PhiGpuError):
        phi_b200.load_reads(broken)
    cut = str(tmp_path / "cut.fa.gz")
    with open(cut, "wb") as f:
        f.write(good[:len(good) // 2])
    rd, names = phi_b200.load_reads(cut)
    assert 0 < rd.n_reads < 3000 and np.all(np.diff(rd.read_off.astype(np.int64))[:-1] == 120)
    if os.path.exists(PROBE):                                                # ... exactly what kseq over gzread makes of the cut file
        gfa = str(tmp_path / "g.gfa")
        with open(gfa, "w") as f:
            f.write("S\ta\tACGT\n")
        d = probe(gfa, cut, str(tmp_path / "p.phiarr"))
        assert rd.read_off.tolist() == d["read_off"].tolist() and bytes(rd.read_bases) == bytes(d["read_bases"])


@pytest.mark.parametrize("chunk", [1, 5, 33, 257, 4096])
def test_reads_parsed_in_parallel_chunks_equal_the_serial_parse(tmp_path, monkeypatch, chunk):
    """Big uncompressed / bgzip read files are cut into chunks that guess their first header and are parsed side by side; the chain
    of guesses is checked afterwards and whatever does not fit is parsed again serially.  PHI_HOST_PARSE_CHUNK puts chunk borders
    everywhere in small files: every edge case of this file, plus FASTQ whose quality lines begin with '@', '>' and '+', multi-line
    FASTQ and a malformed record in the middle, must come out as the serial loop (= kseq) reads them."""
    rng = np.random.default_rng(chunk)
    recs = []
    for i in range(400):
        n = int(rng.integers(1, 90))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTNacgt", dtype=np.uint8), n))
        qual = bytes(rng.choice(np.frombuffer(b"@>+I#5", dtype=np.uint8), n))
        style = int(rng.integers(0, 4))
        if style == 0:
            recs.append(b"@q%d c\n" % i + seq + b"\n+\n" + qual + b"\n")
        elif style == 1:                                       # two-line sequence and quality
            h = n // 2
            recs.append(b"@m%d\n" % i + seq[:h] + b"\n" + seq[h:] + b"\n+m%d\n" % i + qual[:h] + b"\n" + qual[h:] + b"\n")
        elif style == 2:
            recs.append(b">f%d\n" % i + seq + b"\n")
        else:
            recs.append(b">g%d x y\n" % i + seq[:n // 3] + b"\r\n" + seq[n // 3:] + b"\n\n")
    corpus = dict(READS_EDGE)
    corpus = {k: v.encode() for k, v in corpus.items()}
    corpus.update(MALFORMED_READS)
    corpus["random.fq"] = b"".join(recs)
    corpus["random_then_malformed.fq"] = b"".join(recs[:200]) + b"@bad\nACGTACGT\n+\nIII\n" + b"".join(recs[200:])
    for fname, data in sorted(corpus.items()):
        path = str(tmp_path / fname)
        with open(path, "wb") as f:
            f.write(data)
        with open(path + ".bgzf.gz", "wb") as f:
            f.write(bgzf_bytes(data, 300))
        monkeypatch.delenv("PHI_HOST_PARSE_CHUNK", raising=False)
        want, want_names = phi_b200.load_reads(path)                            # small file, no chunks: the serial loop
        monkeypatch.setenv("PHI_HOST_PARSE_CHUNK", str(chunk))
        for p in (path, path + ".bgzf.gz"):
            got, names = phi_b200.load_reads(p)
            assert got.read_off.tolist() == want.read_off.tolist() and bytes(got.read_bases) == bytes(want.read_bases) and names == want_names, (fname, p)
    if os.path.exists(PROBE):                                                   # and the serial loop is kseq's
        gfa = str(tmp_path / "g.gfa")
        with open(gfa, "w") as f:
            f.write("S\ta\tACGT\n")
        for fname in ("random.fq", "random_then_malformed.fq"):
            d = probe(gfa, str(tmp_path / fname), str(tmp_path / "p.phiarr"))
            got, _ = phi_b200.load_reads(str(tmp_path / fname))
            assert got.read_off.tolist() == d["read_off"].tolist() and bytes(got.read_bases) == bytes(d["read_bases"]), fname


def test_whole_buffer_inflate_and_zlib_read_the_same(tmp_path, monkeypatch):
    """Single-member gzip and bgzip files go through the library's own whole-buffer deflate decoder (fast_inflate.h; the gzip trailer's
    CRC-32 and size are verified, zlib takes over on any doubt); PHI_HOST_ZLIB_ONLY=1 forces zlib.  Same graph, same reads — for every
    compression level, stored blocks (level 0) included."""
    import zlib
    c = Case("synth_small")
    gfa = str(tmp_path / "g.gfa")
    synth.write_gfa(c.graph, gfa)
    text = open(gfa, "rb").read()
    ro = c.reads.read_off.astype(np.int64)
    fq = b"".join(b"@r%d\n" % i + bytes(c.reads.read_bases[ro[i]:ro[i + 1]]) + b"\n+\n" + b"I" * int(ro[i + 1] - ro[i]) + b"\n" for i in range(c.reads.n_reads))
    want_g = phi_b200.load_gfa(gfa)
    for level in (0, 1, 6, 9):
        for kind in ("gz", "bgzf"):
            pg, pr = str(tmp_path / f"g{level}.{kind}.gfa.gz"), str(tmp_path / f"r{level}.{kind}.fq.gz")
            for path, data in ((pg, text), (pr, fq)):
                with open(path, "wb") as f:
                    f.write(gzip.compress(data, level) if kind == "gz" else bgzf_bytes(data, 20000, level))
            for zlib_only in (False, True):
                if zlib_only:
                    monkeypatch.setenv("PHI_HOST_ZLIB_ONLY", "1")
                else:
                    monkeypatch.delenv("PHI_HOST_ZLIB_ONLY", raising=False)
                g = phi_b200.load_gfa(pg)
                for fld in ("seg_off", "seg_bases", "walk_off", "walk_vtx", "top_order_map"):
                    assert np.array_equal(getattr(g, fld), getattr(want_g, fld)), (level, kind, zlib_only, fld)
                rd, names = phi_b200.load_reads(pr)
                assert np.array_equal(rd.read_off, c.reads.read_off) and np.array_equal(rd.read_bases, c.reads.read_bases), (level, kind, zlib_only)
    # a member with a header name and comment (gzip -N style) and one whose CRC is wrong: the first is read, the second is zlib's to judge
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(fq) + co.flush()
    import struct
    named = b"\x1f\x8b\x08\x18\x00\x00\x00\x00\x00\x03" + b"reads.fq\x00" + b"a comment\x00" + body + struct.pack("<II", zlib.crc32(fq) & 0xffffffff, len(fq))
    p = str(tmp_path / "named.fq.gz")
    with open(p, "wb") as f:
        f.write(named)
    monkeypatch.delenv("PHI_HOST_ZLIB_ONLY", raising=False)
    rd, _ = phi_b200.load_reads(p)
    assert np.array_equal(rd.read_bases, c.reads.read_bases)
    bad = bytearray(named)
    bad[-8] ^= 0xFF                                                          # CRC-32 field
    with open(p, "wb") as f:
        f.write(bytes(bad))
    with pytest.raises(phi_b200.PhiGpuError):                                # gzread reports the data error at the end: so does this loader
        phi_b200.load_reads(p)
