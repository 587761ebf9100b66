"""Deterministic synthetic pangenome graphs + reads of the shapes BASELINE.json names.

Host-side workload generator for bench.py and the tests (numpy only; nothing here is on the product
path).  The graph is an acyclic bubble chain like a chopped Minigraph-Cactus / vcf2gfa graph
(/root/reference/data/chop_graph.sh:3 `--chop 30`, /root/reference/vcf2gfa.py:54 `-m 30`):
backbone pieces between variant sites, each site a bubble with a reference and an alternative allele
(SNV / insertion / deletion / larger SV), nodes chopped to <= `chop` bp, haplotype walks (W-lines)
that pick one allele per site in LD blocks.  Vertex ids increase along the backbone, so the identity
is a valid topological order.
"""
import numpy as np

from ._abi import Graph, Reads

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in (b"AT", b"TA", b"CG", b"GC", b"at", b"ta", b"cg", b"gc"):
    _COMP[_a] = _b


def expand_ranges(starts, lens):
    """Concatenate arange(starts[i], starts[i]+lens[i]) for all i (vectorised)."""
    starts = np.asarray(starts, dtype=np.int64)
    lens = np.asarray(lens, dtype=np.int64)
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64)
    ends = np.cumsum(lens)
    return np.repeat(starts - (ends - lens), lens) + np.arange(total, dtype=np.int64)


class SynthGraph:
    """Graph (flat views) + what is needed to spell any allele vector as a sequence / walk."""

    def __init__(self, graph, piece_first_node, piece_n_nodes, piece_off, piece_len, n_sites, alleles, fa=None, pick=None, block=None):
        self.graph = graph
        self.piece_first_node, self.piece_n_nodes = piece_first_node, piece_n_nodes
        self.piece_off, self.piece_len = piece_off, piece_len
        self.n_sites, self.alleles = n_sites, alleles          # alleles: [n_haps, n_sites], or None for big graphs (use allele_row)
        self.fa, self.pick, self.block = fa, pick, block       # founder alleles [founders, n_sites], founder of (haplotype, LD block), block of a site

    @property
    def n_haps(self):
        return self.pick.shape[0] if self.pick is not None else self.alleles.shape[0]

    def allele_row(self, h):
        """Allele vector of haplotype h (never materialises the whole [n_haps, n_sites] matrix)."""
        if self.alleles is not None:
            return self.alleles[h]
        return self.fa[self.pick[h][self.block], np.arange(self.n_sites)].astype(np.uint8)

    def walk_of_hap_sites(self, h, s_lo, s_hi):
        """The part of haplotype h's walk that runs through the sites [s_lo, s_hi) (from the backbone piece in front of site s_lo to the
        one behind site s_hi - 1): what a GPU that owns a region of the graph needs, without spelling the whole walk."""
        s_lo, s_hi = max(0, int(s_lo)), min(self.n_sites, int(s_hi))
        cols = np.arange(s_lo, s_hi)
        if self.alleles is not None:
            row = self.alleles[h, s_lo:s_hi]
        else:
            row = self.fa[self.pick[h][self.block[s_lo:s_hi]], cols].astype(np.uint8)
        n = s_hi - s_lo
        sel = np.ones(3 * n + 1, dtype=bool)
        sel[1:3 * n:3] = row == 0
        sel[2:3 * n:3] = row == 1
        p = 3 * s_lo + np.nonzero(sel)[0]
        return expand_ranges(self.piece_first_node[p], self.piece_n_nodes[p]).astype(np.uint32)

    def hap_length(self, h):
        """Bases of haplotype h's whole walk (no vertex list is spelled)."""
        row = self.allele_row(h)
        n = self.n_sites
        return int(self.piece_len[0:3 * n + 1:3].sum() + self.piece_len[1:3 * n:3][row == 0].sum() + self.piece_len[2:3 * n:3][row == 1].sum())

    def mosaic_row(self, src_of_site):
        """Allele vector of a mosaic: site s takes the allele of haplotype src_of_site[s]."""
        cols = np.arange(self.n_sites)
        if self.alleles is not None:
            return self.alleles[src_of_site, cols]
        return self.fa[self.pick[src_of_site, self.block], cols].astype(np.uint8)

    def pieces_of(self, allele_vec):
        n = self.n_sites
        sel = np.ones(3 * n + 1, dtype=bool)
        sel[1:3 * n:3] = allele_vec == 0      # ref allele pieces
        sel[2:3 * n:3] = allele_vec == 1      # alt allele pieces
        return np.nonzero(sel)[0]

    def walk_of(self, allele_vec):
        p = self.pieces_of(allele_vec)
        return expand_ranges(self.piece_first_node[p], self.piece_n_nodes[p]).astype(np.uint32)

    def sequence_of(self, allele_vec):
        p = self.pieces_of(allele_vec)
        return self.graph.seg_bases[expand_ranges(self.piece_off[p], self.piece_len[p])]


def make_graph(seed, backbone_len, n_haps, var_spacing=50, chop=30, founders=8, block_sites=400,
               indel_frac=0.10, sv_frac=0.05, max_indel=50, max_sv=10000, snv_only=False,
               lower_frac=0.0, n_frac=0.0, walk_range=None):
    rng = np.random.Generator(np.random.PCG64(seed))
    L = int(backbone_len)
    backbone = _ACGT[rng.integers(0, 4, L)]
    n_est = max(1, int(L / var_spacing * 1.3) + 16)
    kind = rng.random(n_est)
    is_sv = (kind < sv_frac) & (not snv_only)
    is_indel = (kind >= sv_frac) & (kind < sv_frac + indel_frac) & (not snv_only)
    size = np.ones(n_est, dtype=np.int64)
    size[is_indel] = rng.integers(1, max_indel + 1, int(is_indel.sum()))
    size[is_sv] = np.exp(rng.uniform(np.log(50), np.log(max_sv), int(is_sv.sum()))).astype(np.int64)
    is_del = (is_sv | is_indel) & (rng.random(n_est) < 0.5)
    is_ins = (is_sv | is_indel) & ~is_del
    ref_len = np.where(is_ins, 0, size)                       # SNV: 1, deletion: size, insertion: 0
    alt_len = np.where(is_del, 0, size)                       # SNV: 1, insertion: size, deletion: 0
    gap = rng.geometric(1.0 / var_spacing, n_est).astype(np.int64)        # >= 1 backbone base between sites
    pos = np.cumsum(gap) + np.concatenate([[0], np.cumsum(ref_len)[:-1]])
    keep = pos + ref_len < L - 1
    n_sites = int(keep.sum()) if keep.all() else int(np.argmin(keep))
    pos, ref_len, alt_len = pos[:n_sites], ref_len[:n_sites], alt_len[:n_sites]
    is_snv = (ref_len == 1) & (alt_len == 1)

    alt_off = np.concatenate([[0], np.cumsum(alt_len)])
    alt_pool = _ACGT[rng.integers(0, 4, int(alt_off[-1]))]
    snv_at = alt_off[:-1][is_snv]
    refb = backbone[pos[is_snv]]                               # SNV alt must differ from the reference base
    code = np.searchsorted(_ACGT, refb)
    alt_pool[snv_at] = _ACGT[(code + rng.integers(1, 4, len(code))) % 4]

    # pieces, in vertex order: [inter_0, ref_0, alt_0, inter_1, ref_1, alt_1, ..., tail]
    npieces = 3 * n_sites + 1
    prev_end = np.concatenate([[0], pos + ref_len])            # start of inter piece i
    piece_len = np.zeros(npieces, dtype=np.int64)
    piece_src = np.zeros(npieces, dtype=np.int64)              # offset into [backbone | alt_pool]
    piece_len[0:3 * n_sites:3] = pos - prev_end[:-1]
    piece_src[0:3 * n_sites:3] = prev_end[:-1]
    piece_len[1:3 * n_sites:3] = ref_len
    piece_src[1:3 * n_sites:3] = pos
    piece_len[2:3 * n_sites:3] = alt_len
    piece_src[2:3 * n_sites:3] = L + alt_off[:-1]
    piece_len[-1] = L - prev_end[-1]
    piece_src[-1] = prev_end[-1]
    source = np.concatenate([backbone, alt_pool])
    seg_bases = source[expand_ranges(piece_src, piece_len)]
    piece_off = np.concatenate([[0], np.cumsum(piece_len)])[:-1]

    piece_n_nodes = (piece_len + chop - 1) // chop
    piece_first_node = np.concatenate([[0], np.cumsum(piece_n_nodes)])[:-1]
    n_vtx = int(piece_n_nodes.sum())
    node_piece = np.repeat(np.arange(npieces), piece_n_nodes)
    node_idx_in_piece = np.arange(n_vtx) - piece_first_node[node_piece]
    node_len = np.minimum(chop, piece_len[node_piece] - node_idx_in_piece * chop)
    seg_off = np.concatenate([[0], np.cumsum(node_len)]).astype(np.uint64)

    if lower_frac > 0 or n_frac > 0:                            # slow-path exercise: lower-case runs and N runs
        seg_bases = seg_bases.copy()
        nb = len(seg_bases)
        for frac, fn in ((lower_frac, lambda a: a | 0x20), (n_frac, lambda a: np.full_like(a, ord("N")))):
            nrun = int(nb * frac / 20) if frac > 0 else 0
            for s in rng.integers(0, max(1, nb - 40), nrun):
                ln = int(rng.integers(1, 40))
                seg_bases[s:s + ln] = fn(seg_bases[s:s + ln])

    # haplotypes: founders per LD block, allele frequency ~ Beta(0.5, 2)
    freq = rng.beta(0.5, 2.0, n_sites)
    fa = rng.random((founders, n_sites)) < freq
    nblocks = n_sites // block_sites + 1
    block = np.arange(n_sites) // block_sites
    pick = rng.integers(0, founders, (n_haps, nblocks))
    # [n_haps, n_sites]; big graphs (chromosome-scale configs) keep only founders + picks and spell rows on demand
    alleles = fa[pick[:, block], np.arange(n_sites)].astype(np.uint8) if n_haps * n_sites <= (1 << 27) else None

    g = Graph(seg_off, seg_bases, np.zeros(1, dtype=np.uint64), np.zeros(0, dtype=np.uint32),
              np.arange(n_vtx, dtype=np.int32), [])
    sg = SynthGraph(g, piece_first_node, piece_n_nodes, piece_off, piece_len, n_sites, alleles, fa, pick, block)
    lo, hi = walk_range if walk_range is not None else (0, n_haps)      # only these walks are spelled out (multi-GPU shards)
    walks = [sg.walk_of(sg.allele_row(h)) for h in range(lo, hi)]
    g.walk_off = np.concatenate([[0], np.cumsum([len(x) for x in walks])]).astype(np.uint64)
    g.walk_vtx = np.concatenate(walks).astype(np.uint32) if walks else np.zeros(0, dtype=np.uint32)
    g.walk_names = [f"hap{h}.{h}" for h in range(lo, hi)]
    return sg


def make_reads(seed, sg, coverage, read_len=150, sub_err=0.005, len_sigma=0.0, mosaic_block=2000,
               lower_frac=0.0, n_frac=0.0, sample_seed=0):
    """Reads sampled from a held-out mosaic of the graph's haplotypes, both strands, substitution errors."""
    rng = np.random.Generator(np.random.PCG64(seed ^ 0x5EED))
    n_haps = sg.n_haps
    nseg = sg.n_sites // mosaic_block + 1
    src = rng.integers(0, n_haps, nseg)
    mosaic = sg.mosaic_row(src[np.arange(sg.n_sites) // mosaic_block])
    seq = sg.sequence_of(mosaic)
    n = len(seq)
    if sample_seed:                                               # same sample (mosaic), independent read draw
        rng = np.random.Generator(np.random.PCG64((seed ^ 0x5EED) + 7919 * sample_seed))
    n_reads = max(1, int(round(coverage * n / read_len)))
    if len_sigma > 0:
        lens = np.clip(rng.lognormal(np.log(read_len), len_sigma, n_reads).astype(np.int64), 50, n)
    else:
        lens = np.full(n_reads, min(read_len, n), dtype=np.int64)
    starts = (rng.random(n_reads) * (n - lens + 1)).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(lens)])
    bases = seq[expand_ranges(starts, lens)].copy()
    err = np.nonzero(rng.random(len(bases)) < sub_err)[0]
    code = np.searchsorted(_ACGT, bases[err])
    bases[err] = _ACGT[(code + rng.integers(1, 4, len(err))) % 4]
    # reverse-complement half of the reads
    rev = rng.random(n_reads) < 0.5
    rid = np.repeat(np.arange(n_reads), lens)
    within = np.arange(len(bases)) - off[rid]
    src_idx = np.where(rev[rid], off[rid] + lens[rid] - 1 - within, np.arange(len(bases)))
    bases = np.where(rev[rid], _COMP[bases[src_idx]], bases)
    if lower_frac > 0 or n_frac > 0:
        bases = bases.copy()
        m = rng.random(len(bases)) < lower_frac
        bases[m] |= 0x20
        bases[rng.random(len(bases)) < n_frac] = ord("N")
    return Reads(off.astype(np.uint64), bases.astype(np.uint8))


def mosaic_sequence(seed, sg, mosaic_block=2000):
    """The held-out mosaic sample make_reads / make_reads_big draw from (same for both)."""
    rng = np.random.Generator(np.random.PCG64(seed ^ 0x5EED))
    nseg = sg.n_sites // mosaic_block + 1
    src = rng.integers(0, sg.n_haps, nseg)
    return sg.sequence_of(sg.mosaic_row(src[np.arange(sg.n_sites) // mosaic_block]))


READ_BATCH = 1 << 16


def n_reads_big(seq_len, coverage, read_len):
    return max(1, int(round(coverage * seq_len / read_len)))


def make_reads_big(seed, sg, coverage, read_len=150, sub_err=0.005, read_range=None, seq=None):
    """Fixed-length reads for the chromosome-scale configs, generated in independent batches of READ_BATCH reads (batch b has its
    own generator), so that a rank can spell just its own contiguous range [lo, hi) of the global read set and memory stays bounded.
    Same sample, strands and error model as make_reads; the draws differ (make_reads consumes one sequential stream)."""
    if seq is None:
        seq = mosaic_sequence(seed, sg)
    n = len(seq)
    L = min(read_len, n)
    total = n_reads_big(n, coverage, read_len)
    lo, hi = read_range if read_range is not None else (0, total)
    lo, hi = max(0, lo), min(total, hi)
    out = np.empty((max(hi - lo, 0), L), dtype=np.uint8)
    ar = np.arange(L, dtype=np.int64)
    for b in range(lo // READ_BATCH, (hi + READ_BATCH - 1) // READ_BATCH if hi > lo else 0):
        rng = np.random.Generator(np.random.PCG64([seed ^ 0x5EED, 0xB47C4, b]))
        cnt = min(READ_BATCH, total - b * READ_BATCH)
        starts = (rng.random(cnt) * (n - L + 1)).astype(np.int64)
        rev = rng.random(cnt) < 0.5
        bases = seq[starts[:, None] + ar]
        err = rng.random((cnt, L)) < sub_err
        code = np.searchsorted(_ACGT, bases[err])
        bases[err] = _ACGT[(code + rng.integers(1, 4, len(code))) % 4]
        bases[rev] = _COMP[bases[rev][:, ::-1]]
        a, z = max(lo, b * READ_BATCH), min(hi, b * READ_BATCH + cnt)
        out[a - lo:z - lo] = bases[a - b * READ_BATCH:z - b * READ_BATCH]
    off = np.arange(out.shape[0] + 1, dtype=np.uint64) * np.uint64(L)
    return Reads(off, out.reshape(-1))


# ------------------------------------------------------------------ text writers (for the reference CLI)
def write_gfa(graph, path, walks=None):
    so = graph.seg_off.astype(np.int64)
    wo = graph.walk_off.astype(np.int64)
    walks = range(graph.n_walks) if walks is None else walks
    sb = graph.seg_bases.tobytes()
    with open(path, "w") as f:
        f.write("H\tVN:Z:1.1\n")
        f.write("".join(f"S\ts{v + 1}\t{sb[so[v]:so[v + 1]].decode('latin-1')}\n" for v in range(graph.n_vtx)))
        edges = set()
        for h in walks:
            w = graph.walk_vtx[wo[h]:wo[h + 1]].astype(np.int64)
            if len(w) > 1:
                e = np.unique(w[:-1] * (graph.n_vtx + 1) + w[1:])
                edges.update(e.tolist())
        for e in sorted(edges):
            f.write(f"L\ts{e // (graph.n_vtx + 1) + 1}\t+\ts{e % (graph.n_vtx + 1) + 1}\t+\t0M\n")
        for h in walks:
            w = graph.walk_vtx[wo[h]:wo[h + 1]]
            name = graph.walk_names[h] if graph.walk_names else f"hap{h}.{h}"
            sample, _, hap = name.rpartition(".")
            f.write(f"W\t{sample}\t{hap}\tchr\t0\t0\t" + "".join(">s" + s for s in (w.astype(np.int64) + 1).astype(str)) + "\n")


def write_fasta(reads, path):
    ro = reads.read_off.astype(np.int64)
    rb = reads.read_bases.tobytes()
    with open(path, "w") as f:
        f.write("".join(f">r{i}\n{rb[ro[i]:ro[i + 1]].decode('latin-1')}\n" for i in range(reads.n_reads)))
