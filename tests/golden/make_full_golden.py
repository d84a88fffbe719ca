"""Goldens of the BASELINE.json configurations from the UNMODIFIED reference (build container only):

    python tests/golden/make_full_golden.py [full_c1 full_c2 ...]

c1 (MLP, B = 128) exactly as configured; c2 / c3 (vgg19 + deconv32, BatchNorm both) and c4 (resnet18 + ivgg, 3 x 64 x 64) at the
batch sizes of tests/full_cases.py.  Weights, buffers and inputs come from name-seeded generators (tests/full_cases.py), so no
state_dict is stored: the fixtures hold the reference's outputs only (see tests/full_cases.py).  The eval pass runs first (on
the generated running statistics), then one training evaluate + backward (cvae.py:2440-2459).

Harness patches (the reference itself is not modified): matplotlib / seaborn stubs, torch.randn returning the injected noise
for the duration of one evaluate(), and torchvision.models.resnet18 called without pretrained weights (the reference's default
pretrained=True needs a download, conv.py:247-253).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from make_golden import import_reference, injected_noise, t2n  # noqa: E402
import full_cases as fc  # noqa: E402

EXTRA_METHODS = ['softkl-2', 'kl', 'max', 'sum', 'mean', 'std', 'softiws', 'softiws-5', 'softzdist-2', 'mag', 'nstd', 'IYx',
                 'wmse']


def run(cvae_mod, name):
    import torchvision
    orig = torchvision.models.resnet18
    torchvision.models.resnet18 = lambda *a, **k: orig(weights=None)
    try:
        torch.manual_seed(0)
        model = cvae_mod.ClassificationVariationalNetwork(**fc.ctor_kwargs(name))
    finally:
        torchvision.models.resnet18 = orig
    fc.fill_state_(model, chaotic=name in fc.CHAOTIC)
    x, y, eps_tr, eps_te = fc.inputs(name)
    out = {'n_params': np.array(sum(p.numel() for p in model.parameters()))}
    # ---------------- eval / scoring (cvae.py:1629-1687) on the generated running statistics
    model.eval()
    with torch.no_grad():
        with injected_noise(eps_te):
            x_reco, logits, losses, _, mu, log_var, z = model.evaluate(x, z_output=True)
        for k, v in losses.items():
            out['eval.loss.' + k] = t2n(v)
        out['eval.logits'], out['eval.mu'], out['eval.log_var'] = t2n(logits), t2n(mu), t2n(log_var)
        out['eval.x_reco2'] = t2n(x_reco[:2, :2])
        methods = list(dict.fromkeys([m for m in model.ood_methods if not m.startswith('odin')] + EXTRA_METHODS))
        for m, v in model.batch_dist_measures(logits, losses, methods).items():
            out['eval.measure.' + m] = t2n(v)
        out['eval.methods'] = np.array(json.dumps(methods))
        out['eval.predict_methods'] = np.array(json.dumps(list(model.predict_methods)))
        for m in model.predict_methods:
            out['eval.pred.' + m] = t2n(model.predict_after_evaluate(logits, losses, method=m))
    # ---------------- one training evaluate + backward
    model.train()
    model.optimizer.zero_grad()
    with injected_noise(eps_tr):
        x_reco, logits, losses, measures, mu, log_var, z = model.evaluate(x, y, with_beta=True, z_output=True)
    for k, v in losses.items():
        out['train.loss.' + k] = t2n(v)
    out['train.logits'], out['train.mu'], out['train.log_var'] = t2n(logits), t2n(mu), t2n(log_var)
    out['train.x_reco2'] = t2n(x_reco[:2, :2])
    for k, v in measures.items():
        out['train.measure.' + k] = np.array(float(v))
    losses['total'].mean().backward()
    for k, p in model.named_parameters():
        if p.grad is not None:
            nrm, proj = fc.project(k, t2n(p.grad))
            out['train.gnorm.' + k], out['train.gproj.' + k] = np.array(nrm), proj
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print(name, 'ok', '%.1f KB' % (os.path.getsize(path) / 1024), 'params', int(out['n_params']),
          'train total', float(out['train.loss.total'].mean()))


if __name__ == '__main__':
    os.chdir('/tmp')
    mod = import_reference()
    torch.set_num_threads(int(os.environ.get('THREADS', '8')))
    for name in (sys.argv[1:] or list(fc.CASES)):
        run(mod, name)
