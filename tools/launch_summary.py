#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel totals and the per-launch list of one step."""
import collections, csv, sys

def us(r):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    return v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)

def main(path, detail=False):
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if not l.startswith('==')]))
    ad = [i for i, r in enumerate(rows) if 'adam_kernel' in r['Kernel Name']]
    if len(ad) >= 3:
        a, b = ad[-3], ad[-2]
    else:
        a, b = -1, len(rows) - 1
    step = rows[a + 1:b + 1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in step:
        agg[r['Kernel Name']][0] += 1; agg[r['Kernel Name']][1] += us(r)
    tot = sum(v[1] for v in agg.values())
    print(f'one step: {len(step)} launches, {tot:.1f} us of kernel time')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f'{v[1]:10.1f} us {v[0]:5d}  {100 * v[1] / tot:5.1f}%  {k[:100]}')
    if detail:
        for r in step:
            n = r['Kernel Name']
            if 'conv_' in n or 'gemm' in n:
                print(f"{us(r):9.1f} us grid={r['Grid Size']:>14} {n[:40]}")

if __name__ == '__main__':
    main(sys.argv[1], len(sys.argv) > 2)
