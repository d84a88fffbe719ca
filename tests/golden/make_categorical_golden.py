"""Golden fixtures for output_distribution='categorical' (256-way per-pixel cross-entropy head: losses.py:30-49,
cvae.py:654-660, 683, 776) from the UNMODIFIED reference, same recipe and keys as make_golden.py; named cat_* so the
generic fixture loops (which also pin the oracle's Gaussian-output restatement) leave them to tests/test_gpu_categorical.py.

    python tests/golden/make_categorical_golden.py        # build container only
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, run_case  # noqa: E402

CASES = {
    'cat_mlp_cvae': dict(
        input_shape=(1, 4, 4), num_labels=4, type='cvae', encoder=[24], decoder=[24], classifier=[],
        latent_dim=6, latent_sampling=2, test_latent_sampling=3, gamma=0, beta=1.0, output_distribution='categorical',
        output_activation='linear', sigma={'value': 1.0},
        prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 21}),
    'cat_conv_cvae': dict(
        input_shape=(3, 8, 8), num_labels=4, type='cvae', features='[x3+1]8-8:2-16:2', upsampler='[x3+1]16x2+0-8:2++1-8:2++1-!3x3+1',
        batch_norm=False, encoder=[], decoder=[], classifier=[], latent_dim=8, latent_sampling=2, test_latent_sampling=2,
        gamma=0, beta=1.0, output_distribution='categorical', output_activation='linear', sigma={'value': 1.0},
        prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 22}),
}

if __name__ == '__main__':
    os.chdir('/tmp')
    mod = import_reference()
    torch.set_num_threads(1)
    import numpy as np
    for name, kw in CASES.items():
        run_case(mod, name, kw)
        # the (L+1, B, 256, *shape) logits are megabytes: keep their arg-max image only (what wmse is computed from)
        path = os.path.join(HERE, name + '.npz')
        d = dict(np.load(path))
        for mode in ('train', 'eval'):
            xr = d.pop(mode + '.x_reco')
            d[mode + '.x_reco_argmax'] = xr.argmax(2).astype(np.uint8)
        np.savez_compressed(path, **d)
        print(name, '->', '%.1f KB' % (os.path.getsize(path) / 1024))
