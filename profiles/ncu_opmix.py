#!/usr/bin/env python
"""Opcode mix and top stall sites from an `ncu --page source --csv` export."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cnt = collections.Counter()
tot = 0
samples = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ix['Instructions Executed']])
    except ValueError:
        continue
    toks = r[ix['Source']].split()
    op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
    cnt[op.split('.')[0]] += n
    tot += n
    samples.append((int(r[ix['# Samples']] or 0), n, r[ix['Source']].strip()))
print("warp instructions executed:", tot)
for k, v in cnt.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"  {k:12s} {v:12d} {100 * v / tot:5.1f}%")
print("top stall-sample sites:")
ts = sum(s for s, _, _ in samples)
for s, n, src in sorted(samples, reverse=True)[:15]:
    print(f"  {100 * s / ts:5.1f}%  exec={n:9d}  {src[:90]}")
