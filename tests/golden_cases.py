"""Loader for tests/golden/*.npz (written by tests/golden/make_golden.py from the unmodified reference)."""
import json
import os

import numpy as np

from phi_b200 import _abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL = ["toy_k3_w2", "toy_defaults", "synth_small", "synth_dirty", "synth_k15_w10", "synth_k32_w1",
         "synth_k21_w40_long", "synth_unchopped", "synth_repeats"]
MHC = ["mhc4", "mhc4_N75", "mhc4_lower"]
# shapes of BASELINE.json configs[2] (15 kb reads) and configs[3] (200 haplotypes, fractional threshold): they pin the oracle on those
# shapes against the unmodified reference (the GPU parity tests of the same shapes compare with the oracle)
SHAPES = ["shape_long_reads_15kb", "shape_200_haplotypes"]
# k > 32 (the library's byte-wise path, round 2): pins the ORACLE against the unmodified reference for long k-mers, dirty bytes included;
# the GPU parity tests of k = 33 / 64 / 101 compare with the oracle (tests/test_gpu_parity.py)
LONG_K = ["synth_k33", "synth_k64_dirty", "synth_k101_w9"]


class Case:
    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.z = {k: z[k] for k in z.files}
        self.meta = json.loads(bytes(self.z["meta"]).decode())
        self.digests = json.loads(bytes(self.z["digests"]).decode())
        base = self.z
        if "seg_off" not in self.z:                 # derived from the README case: rebuild the inputs
            b = np.load(os.path.join(GOLDEN, "mhc4.npz"))
            base = {k: b[k] for k in b.files}
        self.graph = _abi.Graph(base["seg_off"], base["seg_bases"], base["walk_off"], base["walk_vtx"], base["top_order_map"],
                                self.meta["walk_names"])
        rb = base["read_bases"].copy()
        ro = base["read_off"].astype(np.int64)
        if name == "mhc4_N75":                       # base 75 (1-based) of every read -> 'N'  (SURVEY.md §9 rule 1)
            ln = np.diff(ro)
            rb[ro[:-1][ln > 74] + 74] = ord("N")
        elif name == "mhc4_lower":
            rb = rb | 0x20
        self.reads = _abi.Reads(base["read_off"], rb)
        self.k, self.w, self.T = self.meta["k"], self.meta["w"], self.meta["T"]

    def expected_anchors(self):
        """[(rank, walk, j, [vertices] or None)] — vertex lists of single-vertex anchors are not in the model dump."""
        z = self.z
        off = z["anchor_off"].astype(np.int64)
        return [(int(z["anchor_rank"][a]), int(z["anchor_walk"][a]), int(z["anchor_j"][a]),
                 z["anchor_vtx"][off[a]:off[a + 1]].tolist() or None) for a in range(len(z["anchor_rank"]))]


def check_against_golden(case, res):
    """res: IndexResultPy from the oracle or the GPU path."""
    m = case.meta
    assert res.count_sp_r == m["count_sp_r"]
    assert res.minimizers_per_walk.tolist() == m["minimizers_per_walk"]
    assert res.anchors_per_walk.tolist() == m["anchors_per_walk"]
    if "spectrum" in case.z:
        assert np.array_equal(res.spectrum, case.z["spectrum"])
    import phi_io
    assert phi_io.sha(res.spectrum) == case.digests["spectrum"]
    if res.count_sp_r:
        assert "%.2f" % (np.float32(res.n_filtered) / np.float32(res.count_sp_r) * 100) == m["filtered_pct"]
    exp = case.expected_anchors()
    assert res.n_anchors == len(exp)
    assert np.array_equal(res.anchor_rank, case.z["anchor_rank"])
    assert np.array_equal(res.anchor_walk, case.z["anchor_walk"])
    # j index = running count inside (rank, walk)
    key = res.anchor_rank.astype(np.int64) * (res.n_walks + 1) + res.anchor_walk
    first = np.r_[True, key[1:] != key[:-1]]
    idx = np.arange(len(key))
    j = idx - np.maximum.accumulate(np.where(first, idx, 0))
    assert np.array_equal(j, case.z["anchor_j"])
    # vertex lists: multi-vertex anchors must match the model dump exactly
    nv = np.diff(res.anchor_off.astype(np.int64))
    gnv = np.diff(case.z["anchor_off"].astype(np.int64))
    multi = nv >= 2
    assert np.array_equal(multi, gnv >= 2)
    assert np.array_equal(nv[multi], gnv[multi])
    got = np.concatenate([res.anchor_vtx[res.anchor_off[a]:res.anchor_off[a + 1]] for a in np.nonzero(multi)[0]]) if multi.any() else np.zeros(0)
    assert np.array_equal(got, case.z["anchor_vtx"])


def assert_same_result(a, b):
    """Field-by-field equality of two IndexResultPy (oracle vs GPU)."""
    assert a.count_sp_r == b.count_sp_r
    assert a.n_filtered == b.n_filtered
    for f in ("spectrum", "anchor_rank", "anchor_walk", "anchor_off", "anchor_vtx", "minimizers_per_walk", "anchors_per_walk"):
        x, y = getattr(a, f), getattr(b, f)
        assert x.shape == y.shape and np.array_equal(x, y), f
    assert a.path_kmer_positions == b.path_kmer_positions
    assert a.read_kmer_positions == b.read_kmer_positions
    assert a.path_minimizers_emitted == b.path_minimizers_emitted
    assert a.path_hits == b.path_hits
