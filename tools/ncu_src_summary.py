"""Summarise `ncu --page source --csv` output: samples per code region (split at marker opcodes) and the hottest instructions.
usage: python tools/ncu_src_summary.py file.csv [marker-regex] [top]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
marker = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"CREDUX\.MAX|WARPSYNC|BRA ")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[2:]:
    if len(r) <= iexec:
        continue
    try:
        data.append((r[ia], r[isrc], int(r[isamp] or 0), int(r[iexec] or 0)))
    except ValueError:
        pass
tot = sum(d[2] for d in data)
print(f"{len(data)} instructions, {tot} samples")
seg_start, seg_s, seg_n, seg_e = 0, 0, 0, 0
for i, (a, src, s, e) in enumerate(data):
    seg_s += s; seg_n += 1; seg_e += e
    if marker.search(src) or i == len(data) - 1:
        if seg_s > tot * 0.004:
            print(f"  [{seg_start:5d}..{i:5d}] {seg_n:5d} instrs  {100.0*seg_s/tot:5.1f}% samples  exec/instr {seg_e/max(seg_n,1):9.0f}   ends: {src[:60]}")
        seg_start, seg_s, seg_n, seg_e = i + 1, 0, 0, 0
print("hottest:")
for a, src, s, e in sorted(data, key=lambda d: -d[2])[:top]:
    print(f"  {100.0*s/tot:5.2f}%  {a[-6:]}  {src[:90]}")
