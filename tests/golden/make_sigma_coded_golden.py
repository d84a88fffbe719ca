"""Golden fixture for Sigma coded by the encoder (train.py:143-161 `--sigma coded`; layers.py:297-298, 398-399;
cvae.py:631-634): one log sigma per sample from the encoder's sigma head.  Same recipe and keys as make_golden.py, from the
UNMODIFIED reference; named sig_* (the oracle's network restatement has no sigma head), used by
tests/test_gpu_model.py::test_sigma_coded_matches_reference.

    python tests/golden/make_sigma_coded_golden.py        # build container only
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, run_case  # noqa: E402

CASES = {
    'sig_mlp_cvae_coded': dict(
        input_shape=(1, 8, 8), num_labels=5, type='cvae', encoder=[32, 16], decoder=[16, 32], classifier=[],
        latent_dim=8, latent_sampling=3, test_latent_sampling=4, gamma=0, beta=1.0, output_activation='sigmoid',
        sigma={'input_dim': [1, 8, 8], 'sdim': 1},
        prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 31}),
    # constant sigma decaying towards reach * rmse of every training batch (layers.py:146-168, cvae.py:768-771)
    'sig_mlp_cvae_decay': dict(
        input_shape=(1, 8, 8), num_labels=5, type='cvae', encoder=[32, 16], decoder=[16, 32], classifier=[],
        latent_dim=8, latent_sampling=3, test_latent_sampling=4, gamma=0, beta=1.0, output_activation='sigmoid',
        sigma={'value': 0.5, 'decay': 0.1, 'reach': 2.0},
        prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 32}),
    # sigma^2 := the sample's own mean squared error (train.py --sigma rmse; cvae.py:662-670), the parameter tracks reach * rmse
    'sig_mlp_cvae_rmse': dict(
        input_shape=(1, 8, 8), num_labels=5, type='cvae', encoder=[32, 16], decoder=[16, 32], classifier=[],
        latent_dim=8, latent_sampling=3, test_latent_sampling=4, gamma=0, beta=1.0, output_activation='sigmoid',
        sigma={'is_rmse': True},
        prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 33}),
}

if __name__ == '__main__':
    os.chdir('/tmp')
    mod = import_reference()
    torch.set_num_threads(1)
    for name, kw in CASES.items():
        run_case(mod, name, kw)
