"""Golden vectors for misclassification_detection_rates (cvae.py:1913-2079) from the UNMODIFIED reference.

The method reads its inputs through the result registry (available_results) and a `record-<set>.pth` file; here the two
lookups are pointed at an in-memory LossRecorder filled with the reference model's own per-class losses of a random
batch, and the unmodified method body runs from there: predictions per predict method, correct / missed split, scores per
misclassification method, ROC table, precision at the kept thresholds.  Saved: the recorder tensors and everything the
method wrote to `model.testing[epoch]`.

    python tests/golden/make_misclass_golden.py        # build container only
"""
import collections
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, injected_noise, t2n  # noqa: E402

CTOR = dict(input_shape=(1, 8, 8), num_labels=5, type='cvae', encoder=[32, 16], decoder=[16, 32], classifier=[],
            latent_dim=8, latent_sampling=3, test_latent_sampling=4, gamma=0, beta=1.0, output_activation='sigmoid',
            sigma={'value': 0.5}, prior={'init_mean': 2.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 1})


def main():
    os.chdir('/tmp')
    cvae_mod = import_reference()
    torch.set_num_threads(1)
    kw = json.loads(json.dumps(CTOR))
    ctor = dict(kw)
    ctor['input_shape'] = tuple(ctor['input_shape'])
    torch.manual_seed(3)
    model = cvae_mod.ClassificationVariationalNetwork(**ctor)
    model.eval()
    N, C, K, L = 240, kw['num_labels'], kw['latent_dim'], kw['test_latent_sampling']
    g = torch.Generator().manual_seed(8)
    # inputs that carry some class signal, so that predictions are neither all right nor all wrong
    y = torch.randint(0, C, (N,), generator=g)
    protos = torch.rand(C, 1, 8, 8, generator=g)
    x = (0.6 * protos[y] + 0.4 * torch.rand(N, 1, 8, 8, generator=g)).clamp(0, 1)
    eps = torch.randn(L + 1, N, K, generator=g)
    with torch.no_grad(), injected_noise(eps):
        _, logits, losses, _ = model.evaluate(x)
    rec = cvae_mod.LossRecorder(N)
    rec.append_batch(**losses, y_true=y, logits=logits.T)
    out = {'cfg': np.array(json.dumps(kw))}
    for k, v in rec._tensors.items():
        out['rec.' + k] = t2n(v)
    epoch = 7
    avail = {epoch: {'testset': {'where': {'recorders': True, 'json': False},
                                 'recorders': collections.defaultdict(lambda: True)}}}
    cvae_mod.available_results = lambda *a, **k: avail
    cvae_mod.LossRecorder.load = staticmethod(lambda *a, **k: rec)
    model.training_parameters['set'] = 'testset'
    model.saved_dir = '/tmp/none'
    model.misclassification_detection_rates(predict_methods='all', misclass_methods='all', from_where=('recorders',))
    res = model.testing[epoch]
    out['results'] = np.array(json.dumps(res, default=lambda o: o.item() if hasattr(o, 'item') else list(o)))
    np.savez_compressed(os.path.join(HERE, 'misclass_cvae.npz'), **out)
    print('ok', {pm: sorted(k for k in r if isinstance(r[k], dict))[:4] for pm, r in res.items()},
          {pm: r.get('accuracy') for pm, r in res.items()})


if __name__ == '__main__':
    main()
