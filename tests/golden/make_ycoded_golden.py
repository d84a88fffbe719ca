"""Golden fixture for y_is_coded=True (the one-hot label is an encoder input, layers.py:366-369): the TRAIN step only.
The reference's evaluation without labels for such a model fails in the reference itself (cvae.py:451 reshapes the
C-replicated y to x's batch shape), so there is nothing to pin there.  From the UNMODIFIED reference, recipe of
make_golden.py; used by tests/test_gpu_model.py::test_y_is_coded_train_step_matches_reference.

    python tests/golden/make_ycoded_golden.py        # build container only
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, run_case  # noqa: E402

CASES = {
    'ycoded_mlp_cvae': dict(
        input_shape=(1, 8, 8), num_labels=5, type='cvae', y_is_coded=True, encoder=[32, 16], decoder=[16, 32], classifier=[],
        latent_dim=8, latent_sampling=3, test_latent_sampling=4, gamma=0, beta=1.0, output_activation='sigmoid',
        sigma={'value': 0.3}, prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 41}),
}

if __name__ == '__main__':
    os.chdir('/tmp')
    mod = import_reference()
    torch.set_num_threads(1)
    for name, kw in CASES.items():
        run_case(mod, name, kw, eval_part=False)
