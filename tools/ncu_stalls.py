#!/usr/bin/env python
"""Top stall-sample instructions per kernel from an .ncu-rep (source page, SASS).  usage: ncu_stalls.py rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'body': []}; blocks.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = r
    elif cur is not None:
        cur['body'].append(r)
seen = None
for b in blocks:
    if seen is not None and b['body'] == seen:      # the source page repeats each launch
        continue
    seen = b['body']
    h = b['hdr']; si = h.index('Warp Stall Sampling (All Samples)'); so = h.index('Source'); ex = h.index('Instructions Executed')
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    tot = sum(int(r[si]) for r in b['body'] if r[si].isdigit())
    print('==', b['name'][:70], 'samples', tot, 'instrs', len(b['body']))
    agg = {c: 0 for _, c in stall_cols}
    for r in b['body']:
        for i, c in stall_cols:
            if r[i].isdigit(): agg[c] += int(r[i])
    print('   reasons:', ', '.join(f'{c[6:]}={v}' for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for i, r in sorted(enumerate(b['body']), key=lambda t: -(int(t[1][si]) if t[1][si].isdigit() else 0))[:topn]:
        rs = sorted(((int(r[j]) if r[j].isdigit() else 0, c[6:]) for j, c in stall_cols), reverse=True)[:2]
        print(f'   {i:5d} {r[si]:>7} x{r[ex]:>9}  {r[so][:70]:70s} {rs}')
