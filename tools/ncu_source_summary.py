#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export: stall samples per block of SASS instructions.
Usage: python tools/ncu_source_summary.py <src.csv> [instructions per segment]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
seg = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = hi[0]; end = hi[1] - 1 if len(hi) > 1 else len(rows)
hdr = rows[h]; ix = {c: i for i, c in enumerate(hdr)}
body = [r for r in rows[h + 1:end] if r[0].startswith('0x')]
tot = sum(int(r[ix['# Samples']] or 0) for r in body)
print('instrs', len(body), 'samples', tot)
stalls = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
def op(r):
    t = r[ix['Source']].split()
    return (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
for i in range(0, len(body), seg):
    b = body[i:i + seg]
    s = sum(int(r[ix['# Samples']] or 0) for r in b)
    ops = collections.Counter(op(r) for r in b)
    ex = max(int(r[ix['Instructions Executed']] or 0) for r in b)
    st = {k: sum(int(r[ix[k]] or 0) for r in b) for k in stalls}
    st = {k[6:]: v for k, v in st.items() if v > s * 0.05}
    print(i, s, f'{100 * s / max(tot, 1):.1f}%', ex, ops.most_common(4), st)
