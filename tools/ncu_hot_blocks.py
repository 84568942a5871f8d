#!/usr/bin/env python
"""Hot basic blocks of one profiled kernel: ncu_hot_blocks.py report.ncu-rep [top]  (runs of SASS instructions with the same
execution count, ranked by executed instructions, with their share of the stall samples)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 24
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix['Instructions Executed']] or 0) for r in data)
tots = sum(int(r[ix['# Samples']] or 0) for r in data)
print('instructions', tot, 'sass lines', len(data), 'samples', tots)
blocks, cur = [], None
for k, r in enumerate(data):
    n = int(r[ix['Instructions Executed']] or 0); sm = int(r[ix['# Samples']] or 0)
    if cur and cur[2] == n: cur[1] = k; cur[3] += sm
    else:
        if cur: blocks.append(cur)
        cur = [k, k, n, sm]
blocks.append(cur)
big = sorted(blocks, key=lambda b: -max((b[1] - b[0] + 1) * b[2] / max(tot, 1), b[3] / max(tots, 1)))[:top]
for k0, k1, n, sm in sorted(big):
    print(f"{k0:5d}-{k1:5d} len {k1-k0+1:3d} x {n:9d} = {(k1-k0+1)*n/tot*100:5.1f}% inst {sm/tots*100:5.1f}% samples ", data[k0][ix['Source']][:48], '|', data[k1][ix['Source']][:48])
