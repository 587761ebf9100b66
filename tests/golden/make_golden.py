#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/PHI_ref and
oracle/_ref/ref_probe, built by `make -C oracle ref` from /root/reference/src with the recording
Gurobi stub) on the reference's own fixtures and on small synthetic graphs.

Run in the build container only (needs /root/reference).  The .npz files are committed; the GPU
box never needs /root/reference.

Each .npz holds the flat input views exactly as the reference numbers them plus what the reference
produced: per-walk minimizers (hash + vertex list; as arrays for small cases, as sha256 digests
for the README case), per-read hash sets / the spectrum, the final anchors recovered from the model
dump (z_i_j_k creation order + edge variables), the stderr counters and the model-dump digests.
"""
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import phi_io  # noqa: E402
from phi_b200 import synth  # noqa: E402

REF = "/root/reference/test"
PHI_REF = os.path.join(ROOT, "oracle", "_ref", "PHI_ref")
PROBE = os.path.join(ROOT, "oracle", "_ref", "ref_probe")


def file_sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def run_reference(gfa, reads, k, w, T, tmp, q):
    dump = os.path.join(tmp, f"dump_q{q}.txt")
    env = dict(os.environ, PHI_STUB_DUMP=dump)
    cmd = [PHI_REF, "-g", gfa, "-r", reads, "-o", os.path.join(tmp, "out.fa"), "-t", "8", "-k", str(k), "-w", str(w),
           "-T", repr(float(T)), "-q", str(q)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, check=True)
    return dump, p.stderr


def parse_stderr(err, names):
    out = {}
    m = re.search(r"spectrum size: (\d+)", err)
    out["count_sp_r"] = int(m.group(1))
    m = re.search(r"Filtered/Retained Minimizers: ([-\w.]+)/([-\w.]+)%", err)
    out["filtered_pct"], out["retained_pct"] = m.group(1), m.group(2)
    m = re.search(r"([-\w.]+)% Minimizers are in ILP", err)
    out["in_ilp_pct"] = m.group(1)
    sec_min, sec_anc = err.split("Number of Minimizers")[1].split("Number of Anchors")
    out["minimizers_per_walk"] = [int(re.search(rf"^{re.escape(n)} : (\d+)$", sec_min, re.M).group(1)) for n in names]
    out["anchors_per_walk"] = [int(re.search(rf"^{re.escape(n)} : (\d+)$", sec_anc, re.M).group(1)) for n in names]
    return out


def make_case(name, gfa, reads, k=31, w=25, T=1.0, full_minimizers=True, transform=None, store_inputs=True):
    print(f"[golden] {name}", flush=True)
    with tempfile.TemporaryDirectory() as tmp:
        if transform:
            reads = transform(reads, tmp)
        arr = os.path.join(tmp, "probe.phiarr")
        subprocess.run([PROBE, "-g", gfa, "-r", reads, "-o", arr, "-k", str(k), "-w", str(w), "-t", "8"], check=True,
                       capture_output=True)
        d = phi_io.read_phiarr(arr)
        names = bytes(d["walk_names"]).decode().split("\n")[:-1]
        dump1, err1 = run_reference(gfa, reads, k, w, T, tmp, 1)
        z_order, lists, n1 = phi_io.parse_model_dump(dump1)
        sha_q1 = file_sha(dump1)
        dump0, _ = run_reference(gfa, reads, k, w, T, tmp, 0)
        z0, lists0, n0 = phi_io.parse_model_dump(dump0)
        sha_q0 = file_sha(dump0)
        assert z0 == z_order and lists0 == lists, "ILP and IQP dumps disagree on the anchors"
    meta = parse_stderr(err1, names)
    meta.update(k=k, w=w, T=T, model_q1_sha256=sha_q1, model_q0_sha256=sha_q0, model_q1_counts=n1, model_q0_counts=n0,
                walk_names=names)
    # anchors from the dump: (rank, walk, j) in creation order; vertex list, or [] when single-vertex (not recoverable)
    a_rank = np.array([z[0] for z in z_order], dtype=np.int32)
    a_walk = np.array([z[1] for z in z_order], dtype=np.int32)
    a_j = np.array([z[2] for z in z_order], dtype=np.int32)
    a_vtx, a_off = [], [0]
    for z in z_order:
        a_vtx.extend(lists.get(z, []))
        a_off.append(len(a_vtx))
    out = dict(meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
               anchor_rank=a_rank, anchor_walk=a_walk, anchor_j=a_j, anchor_off=np.array(a_off, dtype=np.uint64),
               anchor_vtx=np.array(a_vtx, dtype=np.int32), wm_off=d["wm_off"])
    if store_inputs:      # derived cases (store_inputs=False) rebuild their inputs from the base case inside the test
        out.update(seg_off=d["seg_off"], seg_bases=d["seg_bases"], walk_off=d["walk_off"], walk_vtx=d["walk_vtx"],
                   top_order_map=d["top_order_map"], read_off=d["read_off"], read_bases=d["read_bases"],
                   spectrum=np.unique(d["rh_hash"]))
    if full_minimizers:
        out.update(wm_hash=d["wm_hash"], wm_voff=d["wm_voff"], wm_vtx=d["wm_vtx"])
    digests = dict(wm_hash=phi_io.sha(d["wm_hash"]), wm_voff=phi_io.sha(d["wm_voff"]), wm_vtx=phi_io.sha(d["wm_vtx"]),
                   rh_hash=phi_io.sha(d["rh_hash"]), rh_off=phi_io.sha(d["rh_off"]),
                   spectrum=phi_io.sha(np.unique(d["rh_hash"])), read_bases=phi_io.sha(d["read_bases"]))
    out["digests"] = np.frombuffer(json.dumps(digests).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"   spectrum {meta['count_sp_r']}  anchors {len(z_order)}  filtered {meta['filtered_pct']}%", flush=True)


def reads_transform(fn):
    def tr(path, tmp):
        import gzip
        op = gzip.open if path.endswith(".gz") else open
        with op(path, "rt") as f:
            lines = f.read().split("\n")
        step = 4 if lines[0].startswith("@") else 2
        for i in range(1, len(lines), step):
            lines[i] = fn(lines[i])
        out = os.path.join(tmp, "reads_tr." + ("fq" if step == 4 else "fa"))
        with open(out, "w") as f:
            f.write("\n".join(lines))
        return out
    return tr


def synth_case(name, seed, backbone, haps, cov, k=31, w=25, T=1.0, read_len=150, **gk):
    with tempfile.TemporaryDirectory() as tmp:
        rk = {a: gk.pop(a) for a in list(gk) if a.startswith("r_")}
        sg = synth.make_graph(seed, backbone, haps, **gk)
        rd = synth.make_reads(seed, sg, cov, read_len=read_len, **{a[2:]: v for a, v in rk.items()})
        gfa, fa = os.path.join(tmp, "g.gfa"), os.path.join(tmp, "r.fa")
        synth.write_gfa(sg.graph, gfa)
        synth.write_fasta(rd, fa)
        make_case(name, gfa, fa, k, w, T)


if __name__ == "__main__":
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    ONLY = set(sys.argv[1:])                                   # `make_golden.py name ...` regenerates the named cases only
    if ONLY:
        _make_case, _synth_case = make_case, synth_case
        make_case = lambda name, *a, **k: _make_case(name, *a, **k) if name in ONLY else None      # noqa: E731
        synth_case = lambda name, *a, **k: _synth_case(name, *a, **k) if name in ONLY else None    # noqa: E731
    # k > 32 (round 2: the library's byte-wise compare-and-hash path; the reference's string code takes any k, ILP_index.cpp:390-394)
    synth_case("synth_k33", 21, 40000, 5, 3.0, k=33, w=25)
    synth_case("synth_k64_dirty", 22, 40000, 4, 3.0, k=64, w=11, T=0.8, lower_frac=0.02, n_frac=0.0008, r_lower_frac=0.02, r_n_frac=0.0008)
    synth_case("synth_k101_w9", 23, 50000, 4, 4.0, k=101, w=9, read_len=400, T=0.75)
    make_case("toy_k3_w2", f"{REF}/test.gfa", f"{REF}/read.fa", k=3, w=2)
    make_case("toy_defaults", f"{REF}/test.gfa", f"{REF}/read.fa")                       # all walks < 55 bp: nothing
    synth_case("synth_small", 11, 60000, 5, 3.0)
    synth_case("synth_dirty", 12, 50000, 6, 3.0, T=0.5, lower_frac=0.02, n_frac=0.004, r_lower_frac=0.02, r_n_frac=0.004)
    synth_case("synth_k15_w10", 13, 40000, 4, 2.0, k=15, w=10, T=0.75)
    synth_case("synth_k32_w1", 14, 20000, 3, 2.0, k=32, w=1)
    synth_case("synth_k21_w40_long", 15, 60000, 4, 4.0, k=21, w=40, read_len=3000, r_len_sigma=0.3, var_spacing=25, chop=8)
    synth_case("synth_unchopped", 16, 80000, 7, 3.0, chop=100000, T=0.9)
    synth_case("synth_repeats", 17, 30000, 5, 5.0, T=2.0, var_spacing=200)               # T > 1: nothing filtered -> multi-occurrence order
    # shapes of BASELINE.json configs[2] / configs[3] at a size the reference finishes in seconds (oracle pin only: tests/test_oracle_golden.py)
    synth_case("shape_long_reads_15kb", 0x50484931 + 2, 60000, 6, 6.0, read_len=15000, r_len_sigma=0.2, r_sub_err=0.01)
    synth_case("shape_200_haplotypes", 0x50484931 + 3, 12000, 200, 10.0, T=0.335, sv_frac=0.0, max_indel=20, founders=24, block_sites=60)
    make_case("mhc4", f"{REF}/MHC_4.gfa.gz", f"{REF}/CHM13_reads.fq.gz", full_minimizers=False)
    make_case("mhc4_N75", f"{REF}/MHC_4.gfa.gz", f"{REF}/CHM13_reads.fq.gz", full_minimizers=False, store_inputs=False,
              transform=reads_transform(lambda s: s[:74] + "N" + s[75:] if len(s) > 74 else s))
    make_case("mhc4_lower", f"{REF}/MHC_4.gfa.gz", f"{REF}/CHM13_reads.fq.gz", full_minimizers=False, store_inputs=False,
              transform=reads_transform(str.lower))
