import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, './tests')
import numpy as np, torch
import __graft_entry__ as g
pkg = g.build()
import test_gpu_multistep as t
for name in ['multistep_conv_cvae_bn', 'multistep_mlp_cvae']:
    o = t._drive(pkg, name)
    d = o['d']
    for i, e in enumerate(o['errs']):
        print(name, 'step', i, {k: round(v, 4) for k, v in e.items()})
    for tag in ('eval_a', 'eval_b'):
        print(tag, {k: round(v, 4) for k, v in o[tag][0].items()})
    print('twin', o['twin'])
    for i, m in enumerate(o['meas']):
        print('meas', i, {k: (round(a, 4), round(b, 4)) for k, (a, b) in m.items()})
    tr = []
    for k, v in o['sd'].items():
        want, start = d['sd_after.' + k].astype(np.float64), d['sd.' + k].astype(np.float64)
        if want.dtype.kind != 'f' or 'num_batches' in k: continue
        tv = np.linalg.norm(want - start)
        if tv > 1e-6 * max(1.0, np.linalg.norm(start)): tr.append((round(float(np.linalg.norm(v - want) / tv), 3), k))
    print(sorted(tr)[-8:])
