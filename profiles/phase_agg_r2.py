#!/usr/bin/env python
"""Phase-level instruction / stall shares of a sketch kernel from `ncu --page source --csv --print-source cuda,sass` (line ranges of the
round-2 sources).  usage: phase_agg_r2.py src.csv walk|read"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
kind = sys.argv[2]
cur, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = dict(zip(hdr[4:], r[4:]))
        def num(k):
            try:
                return float(d.get(k, "0").replace(",", "") or 0)
            except ValueError:
                return 0.0
        out.append((cur, int(r[0]), num("Instructions Executed"), num("Warp Stall Sampling (All Samples)"), num("Thread Instructions Executed")))
tot = sum(o[2] for o in out); tots = sum(o[3] for o in out); tott = sum(o[4] for o in out)
print(f"total warp instructions {tot:.0f}, thread instructions {tott:.0f} (avg active lanes {tott / tot:.1f}), stall samples {tots:.0f}")
T = "sketch_tile.cuh"; D = "device_common.cuh"; C = "sketch_common.cuh"
K = "walk_sketch.cu" if kind == "walk" else "read_sketch.cu"
phases = [("carve / layout / bounds", T, 76, 113), ("upcase + stage_chunk (pack, dirty mask)", T, 114, 146), ("extract_kmer / extract8", T, 147, 161),
          ("general (shared-memory) core", T, 162, 400), ("window_minima (cross-lane combine)", T, 401, 436), ("fast_runs: geometry", T, 437, 452),
          ("fast_runs: canonical k-mers rolled", T, 453, 469), ("fast_runs: in-lane prefix / suffix minima", T, 470, 485), ("fast_runs: dispatch on r", T, 486, 499),
          ("fast_runs: validity, changed, starts", T, 500, 522), ("fast_runs: compaction of run starts", T, 523, 560),
          ("murmur + fmix", D, 1, 87), ("toupper / is_acgt / code2 / comp_byte", D, 88, 99), ("rev2 / revcomp2", D, 100, 111), ("ascii expansion + hash_packed_kmer", D, 112, 136),
          ("block_scan2 / table_insert / load8", C, 1, 100)]
if kind == "walk":
    phases += [("spectrum_probe", K, 14, 30), ("anchor_slow / step_of", K, 31, 57), ("tile body: runs -> hash -> emit flags", K, 58, 92), ("tile body: probe + anchor size", K, 93, 110),
               ("tile body: scan + segment + hit / probe record write", K, 111, 147), ("kernel: tile record, step table", K, 148, 187), ("kernel: gather bases through the steps", K, 188, 218),
               ("kernel: dispatch", K, 219, 235)]
else:
    phases += [("tile body: runs -> hash -> insert", K, 13, 53), ("kernel: set-up + bulk copy issue + first load", K, 54, 95), ("kernel: boundaries -> window masks", K, 96, 127),
               ("kernel: wait for the bulk copy + stage bases", K, 128, 150), ("kernel: dispatch", K, 151, 160)]
acc = 0
for name, f, a, b in phases:
    i = sum(o[2] for o in out if o[0] == f and a <= o[1] <= b)
    s = sum(o[3] for o in out if o[0] == f and a <= o[1] <= b)
    acc += i
    print(f"{100 * i / tot:6.2f}% inst {100 * s / tots:6.2f}% stall  {name}")
other = {}
for o in out:
    if o[0] not in (T, D, C, K):
        other[o[0]] = other.get(o[0], 0) + o[2]
for k, v in sorted(other.items(), key=lambda x: -x[1])[:6]:
    acc += v
    print(f"{100 * v / tot:6.2f}% inst                {k}")
print(f"covered {100 * acc / tot:.1f}%")
