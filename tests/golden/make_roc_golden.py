"""Generates tests/golden/roc_cases.npz by running the UNMODIFIED reference utils/roc_curves.py (imported from
/root/reference) on seeded score vectors.  Run in the build container:  python tests/golden/make_roc_golden.py"""
import os
import sys
import types

import numpy as np

for name in ('matplotlib', 'matplotlib.pyplot', 'seaborn'):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, '/root/reference')
from utils.roc_curves import roc_curve      # noqa: E402

rng = np.random.default_rng(0)
kept = [pc / 100 for pc in range(90, 100)]
cases = {}
specs = [
    ('sep', rng.normal(2.0, 1.0, 2000), rng.normal(0.0, 1.0, 1500)),
    ('overlap', rng.normal(0.3, 1.0, 1000), rng.normal(0.0, 1.5, 3000)),
    ('ties', np.round(rng.normal(1.0, 1.0, 800), 1), np.round(rng.normal(0.0, 1.0, 700), 1)),
    ('small', rng.normal(1.0, 1.0, 37), rng.normal(0.0, 1.0, 23)),
    ('worse', rng.normal(-0.5, 1.0, 500), rng.normal(0.5, 1.0, 500)),
]
out = {}
for name, ins, outs in specs:
    ins, outs = ins.astype(np.float64), outs.astype(np.float64)
    out[f'{name}.ins'], out[f'{name}.outs'] = ins, outs
    for tag, kw in (('one', {}), ('2s', {'two_sided': 'around-mean'}), ('a11', {'two_sided': (1, 1)}), ('a41', {'two_sided': (4, 1)}),
                    ('lower', {'ins_are_higher': False})):
        auc, fpr, tpr, thr = roc_curve(ins, outs, *kept, **kw)
        out[f'{name}.{tag}.auc'] = np.float64(auc)
        out[f'{name}.{tag}.fpr'], out[f'{name}.{tag}.tpr'] = np.asarray(fpr), np.asarray(tpr)
        out[f'{name}.{tag}.low'], out[f'{name}.{tag}.up'] = np.asarray(thr['low']), np.asarray(thr['up'])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'roc_cases.npz'), **out)
print('written', len(out), 'arrays')
