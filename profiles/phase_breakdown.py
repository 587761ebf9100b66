#!/usr/bin/env python
"""Group an ncu cuda,sass source export of a sketch kernel into pipeline phases by source line range."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
cur_file, hdr = None, None
acc = {}
def phase(f, ln, src):
    if f == "device_common.cuh":
        if ln <= 50: return "hash (murmur)"
        if 52 <= ln <= 64: return "stage (upcase/code/acgt)"
        if 66 <= ln <= 78: return "rev2/revcomp (canon + hash)"
        if 79 <= ln <= 110: return "hash (ascii expand)"
        return "misc"
    if f == "sketch_tile.cuh":
        if ln <= 91: return "setup"
        if 92 <= ln <= 107: return "stage (upcase/code/acgt)"
        if 108 <= ln <= 116: return "canon"
        if 117 <= ln <= 128: return "runs"
        if 129 <= ln <= 151: return "compare (le/lt)"
        if 152 <= ln <= 168: return "hash (dispatch)"
        if 169 <= ln <= 189: return "canon"
        if 190 <= ln <= 217: return "block minima"
        if 218 <= ln <= 228: return "window argmin"
        if 229 <= ln <= 300: return "runs"
        return "misc"
    if f == "sketch_kernels.cu":
        if 20 <= ln <= 45: return "block scan"
        if 150 <= ln <= 165: return "probe"
        if 166 <= ln <= 185: return "anchor slow"
        if 186 <= ln <= 203: return "setup"
        if 204 <= ln <= 228: return "step load"
        if 229 <= ln <= 264: return "gather bases"
        if 265 <= ln <= 277: return "setup"
        if 278 <= ln <= 293: return "hash (dispatch)"
        if 294 <= ln <= 297: return "probe"
        if 298 <= ln <= 313: return "anchor (search/top order)"
        if 314 <= ln <= 345: return "hit output"
        return "misc"
    return "misc"
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = dict(zip(hdr[4:], r[4:]))
        def num(k):
            try: return float(d.get(k, "0").replace(",", "") or 0)
            except ValueError: return 0.0
        p = phase(cur_file, int(r[0]), r[1])
        a = acc.setdefault(p, [0, 0, 0])
        a[0] += num("Instructions Executed"); a[1] += num("Warp Stall Sampling (All Samples)"); a[2] += num("L1 Wavefronts Shared")
ti = sum(a[0] for a in acc.values()); ts = sum(a[1] for a in acc.values())
for p, a in sorted(acc.items(), key=lambda x: -x[1][0]):
    print(f"{p:32s} inst {100*a[0]/ti:6.2f}%  stall-samples {100*a[1]/ts:6.2f}%  smem wavefronts {a[2]:.3g}")
